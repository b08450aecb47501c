"""Learning-rate schedules for `Trainer.set_lr_schedule` (reference: the example suites' warm-up + decay schedules,
examples/GPT2/optimizers.py and examples/gpt_moe/optimizers/*: linear warm-up over `warmup_steps`, then cosine or linear decay).

A schedule is a plain `step -> lr` function of the 1-based optimizer step, evaluated on the host once per step; the value
reaches the kernels through the executor's `hyper` device tensor, so it also works under CUDA-graph replay.
"""
from __future__ import annotations

import math
from typing import Callable

Schedule = Callable[[int], float]


def constant(lr: float) -> Schedule:
    return lambda step: lr


def warmup_linear_decay(lr: float, warmup_steps: int, total_steps: int, end_ratio: float = 0.0) -> Schedule:
    """0 -> lr linearly over `warmup_steps`, then linearly down to end_ratio * lr at `total_steps` (constant afterwards)."""
    def f(step: int) -> float:
        if warmup_steps > 0 and step <= warmup_steps:
            return lr * step / warmup_steps
        span = max(1, total_steps - warmup_steps)
        frac = min(1.0, max(0.0, (step - warmup_steps) / span))
        return lr * (1.0 - (1.0 - end_ratio) * frac)
    return f


def warmup_cosine(lr: float, warmup_steps: int, total_steps: int, end_ratio: float = 0.1) -> Schedule:
    """0 -> lr linearly over `warmup_steps`, then half a cosine down to end_ratio * lr at `total_steps` (constant afterwards)."""
    def f(step: int) -> float:
        if warmup_steps > 0 and step <= warmup_steps:
            return lr * step / warmup_steps
        span = max(1, total_steps - warmup_steps)
        frac = min(1.0, max(0.0, (step - warmup_steps) / span))
        return lr * (end_ratio + (1.0 - end_ratio) * 0.5 * (1.0 + math.cos(math.pi * frac)))
    return f


def rsqrt_decay(lr: float, warmup_steps: int) -> Schedule:
    """Transformer schedule: linear warm-up, then lr * sqrt(warmup / step)."""
    def f(step: int) -> float:
        w = max(1, warmup_steps)
        return lr * min(step / w, math.sqrt(w / max(1, step)))
    return f


def from_spec(spec: dict, base_lr: float) -> Schedule:
    """Schedule from a plain dict, the form that travels with a graph (`build_training_step(..., schedule={...})` stores it in
    graph.meta["optimizer"], which is part of the serialized graph a client sends to a server):
    {"kind": "warmup_cosine" | "warmup_linear_decay" | "rsqrt_decay" | "constant", "warmup_steps": W, "total_steps": T,
     "end_ratio": r}; `lr` defaults to the optimizer's rate."""
    kind = spec.get("kind", "constant")
    lr = float(spec.get("lr", base_lr))
    if kind == "constant":
        return constant(lr)
    if kind == "warmup_cosine":
        return warmup_cosine(lr, int(spec.get("warmup_steps", 0)), int(spec["total_steps"]), float(spec.get("end_ratio", 0.1)))
    if kind == "warmup_linear_decay":
        return warmup_linear_decay(lr, int(spec.get("warmup_steps", 0)), int(spec["total_steps"]), float(spec.get("end_ratio", 0.0)))
    if kind == "rsqrt_decay":
        return rsqrt_decay(lr, int(spec.get("warmup_steps", 1)))
    raise ValueError(f"schedule kind '{kind}'")
