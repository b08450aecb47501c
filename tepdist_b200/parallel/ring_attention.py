"""Context-parallel attention: the sequence stays split over the ranks of one mesh level, K / V blocks ride a ring.

Rank r of n owns the contiguous token block [r L, (r+1) L) of every sequence (L = S / n) -- its queries never move.  Forward:
n steps; at step t the rank holds the K / V block of rank j = (r - t) mod n, runs the block attention kernel on
(q_r, k_j, v_j) -- causal for the diagonal block j = r, unmasked for j < r, skipped for j > r (causal models) -- and merges
the partial output into the running one by log-sum-exp.  While it computes, the block already travels to rank r + 1
(`batch_isend_irecv`: NCCL on GPUs / NVLink, gloo in the CPU tests).  Backward: the same ring; the fp32 dK / dV accumulator of
a block travels WITH the block and arrives home after n hops; the block kernels use the GLOBAL log-sum-exp and output of
the forward, so every partial is exact (no re-normalisation).

The reference has no counterpart: its long-context story is a token split + whatever XLA SPMD makes of the attention dots
(the VERDICT lists CP / ring attention as the extension to build).  The planner side is the "seq" candidate of
the attention rules (csrc/rules.cc AttentionRule) + the `cp` strategy (parallel/__init__.py); the transform stamps
`cp_levels` / `cp_nums` on the node and the executor routes it here.

Load balance: with contiguous blocks and a causal mask rank r does r + 1 block products (n for the last rank, 1 for the
first), i.e. the step is bounded by the last rank -- the zig-zag block layout that balances it needs the fed sequence
permuted consistently (inputs, labels, position rows) and is not done here.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops.attention import attention_bwd, attention_fwd, attn_merge_, attn_ring_accum_, attn_ring_pack


class RingAttention:
    """One ring = the ranks of one mesh level.  `ranks`: global ranks in ring order, `index`: this rank's position."""

    def __init__(self, group, ranks: List[int], index: int, dry: bool = False):
        self.group, self.ranks, self.index, self.n = group, list(ranks), int(index), len(ranks)
        self.dry = dry                       # timing stand-in: no communication, the local block is reused
        self.bytes_moved = 0

    # ---------------------------------------------------------------- ring plumbing
    def _exchange(self, send: torch.Tensor, recv: torch.Tensor):
        """Start `send` -> next rank, `recv` <- previous rank; returns the requests to wait on."""
        if self.dry or self.n == 1:
            recv.copy_(send)
            return []
        nxt, prv = self.ranks[(self.index + 1) % self.n], self.ranks[(self.index - 1) % self.n]
        self.bytes_moved += send.numel() * send.element_size()
        ops_ = [dist.P2POp(dist.isend, send, nxt, self.group), dist.P2POp(dist.irecv, recv, prv, self.group)]
        return dist.batch_isend_irecv(ops_)

    @staticmethod
    def _wait(reqs) -> None:
        for r in reqs:
            r.wait()

    @staticmethod
    def _work(q: torch.Tensor) -> torch.Tensor:
        """[B,L,H,3,D] scratch whose slot 0 holds q: received K / V blocks are dropped into slots 1 / 2 so that q, k, v share
        strides (what the block kernels' tensor maps expect)."""
        B, L, H, D = q.shape
        w = torch.empty(B, L, H, 3, D, dtype=q.dtype, device=q.device)
        w[:, :, :, 0].copy_(q)
        return w

    # ---------------------------------------------------------------- forward
    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True, scale: Optional[float] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        B, L, H, D = q.shape
        scale = 1.0 / math.sqrt(D) if scale is None else scale
        n, r = self.n, self.index
        o_acc = torch.empty(B, L, H, D, dtype=torch.float32, device=q.device)
        lse_acc, lse_nxt = (torch.empty(B, H, L, dtype=torch.float32, device=q.device) for _ in range(2))
        first = True
        cur = torch.stack((k, v), 3).contiguous()              # [B,L,H,2,D]: the travelling block
        nxt = torch.empty_like(cur)
        work = self._work(q) if n > 1 else None
        for t in range(n):
            j = (r - t) % n
            reqs = self._exchange(cur, nxt) if t < n - 1 else []
            if not (causal and j > r):
                if t == 0:
                    o_j, lse_j = attention_fwd(q, k, v, scale, causal)
                else:
                    work[:, :, :, 1:].copy_(cur)
                    o_j, lse_j = attention_fwd(work[:, :, :, 0], work[:, :, :, 1], work[:, :, :, 2], scale, False)
                attn_merge_(o_acc, lse_acc, lse_nxt, o_j.contiguous(), lse_j.contiguous(), first)   # one kernel per block
                lse_acc, lse_nxt, first = lse_nxt, lse_acc, False
            self._wait(reqs)
            cur, nxt = nxt, cur
        return o_acc.to(q.dtype), lse_acc

    # ---------------------------------------------------------------- backward
    def backward(self, do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor,
                 causal: bool = True, scale: Optional[float] = None, dqkv_out: Optional[torch.Tensor] = None):
        """Returns (dq, dk, dv) as views of `dqkv_out` [B,L,H,3,D]."""
        B, L, H, D = q.shape
        scale = 1.0 / math.sqrt(D) if scale is None else scale
        n, r = self.n, self.index
        if dqkv_out is None:
            dqkv_out = torch.empty(B, L, H, 3, D, dtype=q.dtype, device=q.device)
        do, o, lse = do.contiguous(), o.contiguous(), lse.contiguous()
        cur = torch.stack((k, v), 3).contiguous()
        nxt = torch.empty_like(cur)
        acc = torch.zeros(B, L, H, 2, D, dtype=torch.float32, device=q.device)      # dK / dV of the block that `cur` holds
        acc_in = torch.empty_like(acc)
        dq_acc = torch.zeros(B, L, H, D, dtype=torch.float32, device=q.device)
        work = self._work(q) if n > 1 else None
        part = torch.empty(B, L, H, 3, D, dtype=q.dtype, device=q.device)
        for t in range(n):
            j = (r - t) % n
            reqs = self._exchange(cur, nxt) if t < n - 1 else []
            if not (causal and j > r):
                if t == 0:
                    attention_bwd(do, q, k, v, o, lse, scale, causal, dqkv_out=part)
                else:
                    work[:, :, :, 1:].copy_(cur)
                    attention_bwd(do, work[:, :, :, 0], work[:, :, :, 1], work[:, :, :, 2], o, lse, scale, False, dqkv_out=part)
                attn_ring_accum_(dq_acc, acc, part)
            self._wait(reqs)
            # the accumulator follows its block: after the last step it is one hop from home
            self._wait(self._exchange(acc, acc_in))
            acc, acc_in = acc_in, acc
            cur, nxt = nxt, cur
        if dqkv_out.is_contiguous():
            attn_ring_pack(dq_acc, acc, dqkv_out)
        else:
            dqkv_out[:, :, :, 0].copy_(dq_acc)
            dqkv_out[:, :, :, 1:].copy_(acc)
        return dqkv_out[:, :, :, 0], dqkv_out[:, :, :, 1], dqkv_out[:, :, :, 2]
