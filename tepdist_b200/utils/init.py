"""Deterministic variable initialisation from a serialisable spec.

Variables are never shipped from the client: the client sends shape + this spec and each worker fills
only its shard (reference: xla/rng_distribution_config.proto, pjrt/initializers.{h,cc}).
`init_tensor(..., shard=(attrs, coords))` fills ONLY this rank's shard: the C++ core walks the shard's contiguous runs
of the flattened full tensor (`SliceRuns`) and draws exactly those positions of the counter-based Philox stream
(`_C.philox_fill_shard`), so the values are bit-identical to slicing a full-tensor fill -- which remains available (no
`shard`) as the oracle and as the fallback.
"""
from __future__ import annotations

import hashlib
from typing import Any, Dict, Sequence

import torch


def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{name}:{seed}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def _shard_splits(attrs: Dict[str, Any], coords: Dict[int, int]):
    """(dims, nums, ids) of the recorded splits, in application order (same order runtime.executor.shard_of narrows in)."""
    dims = [int(d) for d in attrs.get("shard_dims", [])]
    nums = [int(n) for n in attrs.get("shard_nums", [])]
    lvls = [int(l) for l in attrs.get("shard_levels", [])]
    return dims, nums, [int(coords.get(l, 0)) for l in lvls]


def shard_shape(shape: Sequence[int], attrs: Dict[str, Any]) -> tuple:
    out = list(shape)
    for d, n in zip(attrs.get("shard_dims", []), attrs.get("shard_nums", [])):
        out[int(d)] //= int(n)
    return tuple(out)


def init_tensor(spec: Dict[str, Any], shape: Sequence[int], seed: int, name: str, shard=None) -> torch.Tensor:
    """Values of the variable `name` with FULL shape `shape`.  shard=None: the whole tensor.  shard=(attrs, coords): only the
    shard described by attrs['shard_dims' / 'shard_nums' / 'shard_levels'] at mesh coordinates `coords`, generated without
    materialising the full tensor."""
    kind = spec.get("kind", "constant")
    shape = tuple(shape)
    if shard is not None and shard[0].get("shard_dims"):
        attrs, coords = shard
        if kind == "constant":
            return torch.full(shard_shape(shape, attrs), float(spec.get("value", 0.0)), dtype=torch.float32)
        try:
            from .. import _C
            import numpy as np
            dims, nums, ids = _shard_splits(attrs, coords)
            levels = [_C.DimStrategy.split(d, n) for d, n in zip(dims, nums)]
            arr = _C.philox_fill_shard(kind, _seed_for(name, int(spec.get("seed", seed))), list(shape), levels, ids,
                                       float(spec.get("mean", 0.0)), float(spec.get("std", 1.0)),
                                       float(spec.get("lo", 0.0)), float(spec.get("hi", 1.0)))
            return torch.from_numpy(np.asarray(arr, dtype=np.float32)).reshape(shard_shape(shape, attrs)).clone()
        except (ImportError, AttributeError):
            from ..runtime.executor import shard_of
            return shard_of(init_tensor(spec, shape, seed, name), attrs, coords).contiguous()
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
