"""tools/run_graph.py -- the counterpart of the reference's run_arbitary_hlo debug binary."""
import json

import pytest

from tepdist_b200.tools import run_graph


@pytest.mark.parametrize("snippet", sorted(run_graph.SNIPPETS))
def test_builtin_snippets_run(snippet, capsys):
    assert run_graph.main(["--snippet", snippet, "--steps", "2", "--device", "cpu"]) == 0
    out = capsys.readouterr().out
    assert out.count("finite True") == 2 and "finite False" not in out


def test_graph_json_round_trip_profile_and_plan(tmp_path, capsys):
    path = str(tmp_path / "g.json")
    assert run_graph.main(["--snippet", "ln_linear", "--device", "cpu", "--dump", path, "--profile", "--plan", "2"]) == 0
    first = capsys.readouterr().out
    assert "layernorm" in first and "S(0/2)" in first and "collectives:" in first
    assert json.load(open(path))["nodes"]
    assert run_graph.main(["--graph", path, "--device", "cpu"]) == 0
    second = capsys.readouterr().out
    line = [l for l in first.splitlines() if l.startswith("step 0")][0]
    assert line in second          # same graph, same seeds -> same numbers
