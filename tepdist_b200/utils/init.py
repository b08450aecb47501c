"""Deterministic variable initialisation from a serialisable spec.

Variables are never shipped from the client: the client sends shape + this spec and each worker fills
only its shard (reference: xla/rng_distribution_config.proto, pjrt/initializers.{h,cc}).
Full-tensor generation here is the oracle; the sharded, bit-identical Philox generator lives in the C++
core (`_C.philox_fill`) and is used by the runtime when a variable is sharded.
"""
from __future__ import annotations

import hashlib
from typing import Any, Dict, Sequence

import torch


def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{name}:{seed}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def init_tensor(spec: Dict[str, Any], shape: Sequence[int], seed: int, name: str) -> torch.Tensor:
    kind = spec.get("kind", "constant")
    shape = tuple(shape)
    if kind == "constant":
        return torch.full(shape, float(spec.get("value", 0.0)), dtype=torch.float32)
    try:
        from .. import _C  # sharded-consistent Philox path
        import numpy as np
        n = 1
        for d in shape:
            n *= d
        arr = _C.philox_fill(kind, _seed_for(name, int(spec.get("seed", seed))), 0, n,
                             float(spec.get("mean", 0.0)), float(spec.get("std", 1.0)),
                             float(spec.get("lo", 0.0)), float(spec.get("hi", 1.0)))
        return torch.from_numpy(np.asarray(arr, dtype=np.float32)).reshape(shape).clone()
    except (ImportError, AttributeError):
        gen = torch.Generator().manual_seed(_seed_for(name, int(spec.get("seed", seed))))
        if kind == "normal":
            return torch.randn(shape, generator=gen) * float(spec.get("std", 1.0)) + float(spec.get("mean", 0.0))
        if kind == "uniform":
            lo, hi = float(spec.get("lo", 0.0)), float(spec.get("hi", 1.0))
            return torch.rand(shape, generator=gen) * (hi - lo) + lo
        if kind == "truncated_normal":
            t = torch.randn(shape, generator=gen).clamp_(-2.0, 2.0)
            return t * float(spec.get("std", 1.0)) + float(spec.get("mean", 0.0))
        raise ValueError(kind)
