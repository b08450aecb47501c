"""Observability: always-on plan log lines, DEBUG artefact dumps, step profiles.

Reference parity: the reference logs the chosen `[Strategy]`, collective counts `num_ar/ag/aa/ds`,
the memory estimate and the schedule dump (auto_parallel.cc:321,391; spmd_transform.cc:1945-1949; execution_plan.cc:
388-394) and, under DEBUG, writes planner artefacts into the working directory (`strategies.txt`, `comm_info.<ord>.txt`,
`sketch_raw.dot`, `dag.dot`, ...).  Here: a `tepdist_b200` logger (level from TEPDIST_LOG, default INFO on rank 0,
WARNING elsewhere) and `dump_plan_artifacts` (enabled by DEBUG=true / TEPDIST_DEBUG=1, directory TEPDIST_DUMP_DIR).
Device-timed per-node profiles come from `Executor.profile` (runtime/executor.py).
"""
from __future__ import annotations

import json
import logging
import os
from typing import Any, Dict, Optional

_LOG: Optional[logging.Logger] = None


def logger() -> logging.Logger:
    global _LOG
    if _LOG is None:
        lg = logging.getLogger("tepdist_b200")
        if not lg.handlers:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("[tepdist %(levelname).1s] %(message)s"))
            lg.addHandler(h)
            lg.propagate = False
        rank0 = int(os.environ.get("RANK", "0")) == 0
        lg.setLevel(getattr(logging, os.environ.get("TEPDIST_LOG", "INFO" if rank0 else "WARNING").upper(), logging.INFO))
        _LOG = lg
    return _LOG


def debug_enabled() -> bool:
    if os.environ.get("TEPDIST_DEBUG", "0") == "1":
        return True
    from .. import config
    return config.debug()


def graph_to_dot(graph, max_nodes: int = 4000) -> str:
    """The step graph as Graphviz (sources omitted; collectives boxed red, backward nodes grey)."""
    from ..ir import COLLECTIVE_OPS, SOURCE_OPS
    out = ["digraph step {", "  rankdir=TB; node [fontsize=9, shape=box, style=rounded];"]
    keep = {n.id for n in graph.nodes[:max_nodes] if n.op not in SOURCE_OPS}
    for n in graph.nodes[:max_nodes]:
        if n.id not in keep:
            continue
        shape = "x".join(str(d) for d in (n.outputs[0].shape if n.outputs else ()))
        color = "red" if n.op in COLLECTIVE_OPS else ("grey50" if n.backward else "black")
        stage = f" s{n.stage}" if n.stage >= 0 else ""
        out.append(f'  n{n.id} [label="{n.op}\\n{n.name}\\n[{shape}]{stage}", color={color}];')
        for v in n.inputs:
            if v.node in keep:
                out.append(f"  n{v.node} -> n{n.id};")
    out.append("}")
    return "\n".join(out)


def log_plan(info: Dict[str, Any], parallelism: str) -> None:
    lg = logger()
    c = info.get("collectives", {}) or {}
    lg.info("[Strategy] %s  comm_bytes=%.3g  solve=%.2fs  subgraphs=%s (distinct %s)", parallelism,
            float(info.get("comm_bytes", 0.0)), float(info.get("solve_seconds", 0.0)), info.get("subgraphs"),
            info.get("distinct_subgraphs"))
    lg.info("[Collectives] num_ar=%d num_ag=%d num_rs=%d num_aa=%d num_ds=%d  grad_buckets=%s", c.get("all_reduce", 0),
            c.get("all_gather", 0), c.get("reduce_scatter", 0), c.get("all_to_all", 0), c.get("dynamic_slice", 0),
            info.get("grad_buckets"))
    if "stages" in info:
        lg.info("[Pipeline] stages=%s micro=%s spmd=%s est_makespan=%.3gs est_bubble=%.3g cut_bytes=%.3g", info.get("stages"),
                info.get("micro"), info.get("spmd"), float(info.get("makespan_est", 0.0)), float(info.get("bubble_est", 0.0)),
                float(info.get("cut_bytes", 0.0)))


def dump_plan_artifacts(graph, info: Dict[str, Any], dirpath: Optional[str] = None, extra: Optional[Dict[str, str]] = None) -> str:
    """Write strategies.txt, comm_info.txt, plan.json, step_graph.dot (+ any `extra` name -> text) and return the directory."""
    d = dirpath or os.environ.get("TEPDIST_DUMP_DIR", os.path.join(os.getcwd(), "tepdist_dump"))
    os.makedirs(d, exist_ok=True)
    if info.get("strategies_txt"):
        with open(os.path.join(d, "strategies.txt"), "w") as f:
            f.write(info["strategies_txt"])
    if info.get("comm_info"):
        with open(os.path.join(d, "comm_info.0.txt"), "w") as f:
            f.write(str(info["comm_info"]))
    if info.get("log"):
        with open(os.path.join(d, "auto_parallel.log"), "w") as f:
            f.write(str(info["log"]))
    with open(os.path.join(d, "plan.json"), "w") as f:
        json.dump({k: v for k, v in info.items() if k not in ("strategies_txt",)}, f, indent=1, default=str)
    with open(os.path.join(d, "step_graph.dot"), "w") as f:
        f.write(graph_to_dot(graph))
    for name, text in (extra or {}).items():
        with open(os.path.join(d, name), "w") as f:
            f.write(text)
    logger().info("planner artefacts written to %s", d)
    return d
