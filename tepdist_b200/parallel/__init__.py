"""Parallel execution: building per-rank executors from a plan (DP / TP / PP / EP)."""
from __future__ import annotations

import torch


def plan_and_build(graph, trainer, strategy, comm_mode, use_cuda_graph, seed):
    """Multi-rank build.  strategy: 'auto' | 'dp' | 'tp' | 'pp...' (see planner).  Round-1 minimum: DP."""
    from ..runtime.executor import Executor
    from .dp import make_nccl_grad_sync
    sync = make_nccl_grad_sync()
    return Executor(graph, trainer.device, seed=seed, grad_sync=sync, use_cuda_graph=use_cuda_graph)
