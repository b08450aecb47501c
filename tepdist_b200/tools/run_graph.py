"""Run an IR graph in isolation -- the counterpart of the reference's `run_arbitary_hlo` debug binary (rpc/run_arbitary_hlo.cc:
hard-coded HLO snippets executed on one GPU through HloRunner).

    python -m tepdist_b200.tools.run_graph --snippet matmul                # built-in snippets: see SNIPPETS
    python -m tepdist_b200.tools.run_graph --graph step_graph.json --steps 3 --profile
    python -m tepdist_b200.tools.run_graph --snippet attention --plan 2     # also print the SPMD plan for 2 devices

`--graph` takes what `Graph.to_json()` / the DEBUG artefact dump (`plan.json`'s graph, utils/trace.py) writes.  Inputs are random
(normal for floats, uniform ids below the `vocab` attribute of their consumer for integers); the graph runs through the same
Executor as training (CUDA kernels on a GPU, torch fallbacks on CPU).  Prints every fetched value's shape / mean / abs-max and,
with --profile, the device time per node kind.
"""
from __future__ import annotations

import argparse
import json
import sys
from typing import Dict

import torch

from ..frontend.builder import GraphBuilder
from ..ir import Graph
from ..runtime.executor import Executor


def _matmul() -> Graph:
    b = GraphBuilder("snippet_matmul", compute_dtype="f32")
    x = b.input("x", (128, 256), "f32")
    w = b.parameter("w", (256, 64), {"kind": "normal", "std": 0.05})
    b.g.outputs = [b.matmul(x, w, name="mm")]
    return b.g


def _layernorm_linear() -> Graph:
    b = GraphBuilder("snippet_ln_linear")
    x = b.input("x", (4, 128, 256), "bf16")
    g_ = b.parameter("g", (256,), {"kind": "constant", "value": 1.0})
    b_ = b.parameter("b", (256,), {"kind": "constant", "value": 0.0})
    w = b.parameter("w", (512, 256), {"kind": "normal", "std": 0.02})
    b.g.outputs = [b.gelu(b.linear(b.layernorm(x, g_, b_), w, name="fc"))]
    return b.g


def _attention() -> Graph:
    from ..models.smoke import build_attention_graph
    return build_attention_graph()


def _conv() -> Graph:
    from ..models.smoke import build_conv_graph
    return build_conv_graph()


def _mlp_step() -> Graph:
    from ..models.smoke import build_mlp_graph
    return build_mlp_graph()


SNIPPETS = {"matmul": _matmul, "ln_linear": _layernorm_linear, "attention": _attention, "conv": _conv, "mlp_step": _mlp_step}


def random_feeds(g: Graph, seed: int = 0) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    feeds = {}
    for n in g.inputs():
        t = n.outputs[0]
        if t.dtype in ("i32", "i64"):
            hi = 2
            for u in g.nodes:
                if any(v.node == n.id for v in u.inputs):
                    hi = max(hi, int(u.attrs.get("vocab", u.attrs.get("classes", 2))))
            feeds[n.name] = torch.randint(0, hi, tuple(t.shape), generator=gen, dtype=torch.int32 if t.dtype == "i32" else torch.int64)
        else:
            feeds[n.name] = torch.randn(tuple(t.shape), generator=gen)
    return feeds


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--snippet", choices=sorted(SNIPPETS))
    src.add_argument("--graph", help="JSON file written by Graph.to_json()")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--profile", action="store_true", help="device time per node kind (one extra eager step)")
    ap.add_argument("--plan", type=int, default=0, metavar="N", help="also plan the graph for N devices and print the strategies")
    ap.add_argument("--dump", help="write the graph as JSON to this file")
    a = ap.parse_args(argv)
    if a.snippet:
        g = SNIPPETS[a.snippet]()
    else:
        d = json.load(open(a.graph))
        g = Graph.from_dict(d.get("graph", d))
    g.validate()
    if a.dump:
        open(a.dump, "w").write(g.to_json())
    print(f"graph '{g.name}': {len(g.nodes)} nodes, {len(g.inputs())} inputs, {len(g.params())} variables, {len(g.outputs)} fetches")
    dev = torch.device(a.device)
    ex = Executor(g, dev, seed=0, use_cuda_graph=False)
    feeds = random_feeds(g)
    for step in range(a.steps):
        outs = ex.step(feeds)
        for v, t in zip(g.outputs, outs):
            tf = t.detach().float()
            print(f"step {step}  {g.nodes[v.node].name}[{v.idx}] shape {tuple(t.shape)} dtype {t.dtype}  mean {float(tf.mean()):+.6g}  "
                  f"absmax {float(tf.abs().max()):.6g}  finite {bool(torch.isfinite(tf).all())}")
    if a.profile:
        prof = ex.profile(feeds, warmup=1)
        for kind, ms in sorted(prof["by_op"].items(), key=lambda kv: -kv[1])[:20]:
            print(f"  {kind:24s} {ms:9.3f} ms")
    if a.plan > 1:
        from ..parallel import plan_spmd
        _, info = plan_spmd(g, a.plan, "auto")
        print(info["strategies_txt"])
        print("collectives:", info["collectives"], " bytes/device:", info["comm_bytes"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
