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


def test_preflight_of_a_multi_gpu_plan_runs_on_cpu(monkeypatch, capsys):
    """bench/preflight_rank0.py: rank 0's executor of an N-way plan, stepped with stand-in collectives (no GPUs, no process group)."""
    import importlib.util
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench", "preflight_rank0.py")
    spec = importlib.util.spec_from_file_location("preflight_rank0", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for strategy in ("auto", "tp"):
        monkeypatch.setattr(sys, "argv", ["preflight_rank0.py", "--model", "tiny", "--gpus", "4", "--steps", "2", "--strategy", strategy])
        mod.main()
        out = capsys.readouterr().out
        assert out.count("partial loss") == 2 and "state_dict:" in out
        assert ("sharded optimizer: True" in out) == (strategy == "auto")
