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
first), i.e. the step is bounded by the last rank.  `zigzag=True` (TEPDIST_CP_ZIGZAG=1; causal only) balances it INSIDE the
op, invisible to the graph: the sequence is seen as 2n chunks, rank r computes for chunks (r, 2n-1-r) -- an early and a late
one.  Its contiguous halves (chunks 2r, 2r+1) are exchanged into that layout before the ring (one point-to-point batch for
q / k / v, one for dO / O backward) and the result is exchanged back.  In the zig-zag layout every ring step costs every rank
the same: the local step is ONE causal product over the rank's block [early; late], a step that holds the block of an
earlier rank j < r is two half-size unmasked products (both query chunks x the early chunk of j), a later rank's block two
others (the late query chunk x both chunks of j) -- n/2 block products per rank instead of up to n - 1/2.  CPU-validated
(gloo, 2 and 4 ranks, against full attention); on GPUs it drives the same block kernels as the contiguous ring but has not
run there yet, hence opt-in.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops.attention import attention_bwd, attention_fwd, attn_merge_, attn_ring_accum_, attn_ring_pack


class RingAttention:
    """One ring = the ranks of one mesh level.  `ranks`: global ranks in ring order, `index`: this rank's position."""

    def __init__(self, group, ranks: List[int], index: int, dry: bool = False, zigzag: Optional[bool] = None):
        import os
        self.group, self.ranks, self.index, self.n = group, list(ranks), int(index), len(ranks)
        self.dry = dry                       # timing stand-in: no communication, the local block is reused
        self.bytes_moved = 0
        self.zigzag = (os.environ.get("TEPDIST_CP_ZIGZAG") == "1") if zigzag is None else bool(zigzag)
        self.block_products = 0.0            # work done by this rank, in units of one full L x L block product (diagnostic)

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
        if self._use_zigzag(q, causal):
            return self._forward_zigzag(q, k, v, scale)
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
                self.block_products += 0.5 if (causal and t == 0) else 1.0
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
        if self._use_zigzag(q, causal):
            return self._backward_zigzag(do, q, k, v, o, lse, scale, dqkv_out)
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
                self.block_products += 0.5 if (causal and t == 0) else 1.0
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

    # ================================================================ zig-zag (load-balanced causal) variant
    def _use_zigzag(self, q: torch.Tensor, causal: bool) -> bool:
        L = q.shape[1]
        return bool(self.zigzag and causal and self.n > 1 and L % 2 == 0 and (not q.is_cuda or (L // 2) % 128 == 0))

    def _owner(self, c: int) -> int:
        """Rank that computes for global chunk c (of 2n) in the zig-zag layout."""
        return c if c < self.n else 2 * self.n - 1 - c

    def _permute(self, x: torch.Tensor, to_zigzag: bool) -> torch.Tensor:
        """[B, L, ...] along dim 1: contiguous layout (chunks 2r, 2r+1) <-> zig-zag layout (chunks r, 2n-1-r); one p2p batch."""
        n, r = self.n, self.index
        l = x.shape[1] // 2
        x = x.contiguous()
        out = torch.empty_like(x)
        mine_c, mine_z = (2 * r, 2 * r + 1), (r, 2 * n - 1 - r)
        have, want = (mine_c, mine_z) if to_zigzag else (mine_z, mine_c)
        holder = (lambda c: c // 2) if to_zigzag else self._owner       # who holds chunk c before the exchange
        target = self._owner if to_zigzag else (lambda c: c // 2)       # who must hold it afterwards
        sends, recvs, keep = [], [], []
        for slot, c in enumerate(have):          # chunk ids ascend on both sides: two messages between one pair stay ordered
            piece = x[:, slot * l:(slot + 1) * l]
            if target(c) == r:
                out[:, want.index(c) * l:(want.index(c) + 1) * l].copy_(piece)
            else:
                buf = piece.contiguous()
                keep.append(buf)
                sends.append((c, dist.P2POp(dist.isend, buf, self.ranks[target(c)], self.group)))
                self.bytes_moved += buf.numel() * buf.element_size()
        for slot, c in enumerate(want):
            if holder(c) != r:
                buf = torch.empty_like(x[:, :l]).contiguous()
                keep.append(buf)
                recvs.append((c, slot, buf, dist.P2POp(dist.irecv, buf, self.ranks[holder(c)], self.group)))
        if self.dry:
            for c, slot, buf, _ in recvs:
                out[:, slot * l:(slot + 1) * l].copy_(x[:, :l])
            return out
        ops_ = [op for _, op in sorted(sends, key=lambda t: t[0])] + [op for _, _, _, op in sorted(recvs, key=lambda t: t[0])]
        if ops_:
            self._wait(dist.batch_isend_irecv(ops_))
        for c, slot, buf, _ in recvs:
            out[:, slot * l:(slot + 1) * l].copy_(buf)
        return out

    @staticmethod
    def _halves(x: torch.Tensor) -> torch.Tensor:
        """[B, L, ...] -> contiguous [2, B, L/2, ...] (the two chunks of the block)."""
        l = x.shape[1] // 2
        return torch.stack((x[:, :l], x[:, l:]), 0).contiguous()

    def _pairs(self, t: int):
        """(query half, key / value half) products of ring step t > 0 in the zig-zag layout."""
        j = (self.index - t) % self.n
        return ((0, 0), (1, 0)) if j < self.index else ((1, 0), (1, 1))

    def _forward_zigzag(self, q, k, v, scale):
        B, L, H, D = q.shape
        n, l, dev = self.n, L // 2, q.device
        qkv = self._permute(torch.stack((q, k, v), 3), True)                   # [B,L,H,3,D] in the zig-zag layout
        qz, kz, vz = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
        o0, lse0 = attention_fwd(qz, kz, vz, scale, True)                      # local step: causal over [early; late]
        self.block_products += 0.5
        o_acc = torch.empty(2, B, l, H, D, dtype=torch.float32, device=dev)
        lse_a = [torch.empty(B, H, l, dtype=torch.float32, device=dev) for _ in range(2)]
        lse_b = [torch.empty_like(t_) for t_ in lse_a]
        oh, lh = self._halves(o0), (lse0[:, :, :l].contiguous(), lse0[:, :, l:].contiguous())
        for h in range(2):
            attn_merge_(o_acc[h], lse_a[h], lse_b[h], oh[h], lh[h], True)
            lse_a[h], lse_b[h] = lse_b[h], lse_a[h]
        cur = self._halves(torch.stack((kz, vz), 3))                           # [2,B,l,H,2,D]: the travelling block
        nxt = torch.empty_like(cur)
        work = [torch.empty(B, l, H, 3, D, dtype=q.dtype, device=dev) for _ in range(2)]
        for h in range(2):
            work[h][:, :, :, 0].copy_(qz[:, h * l:(h + 1) * l])
        reqs = self._exchange(cur, nxt)
        for t in range(1, n):
            self._wait(reqs)
            cur, nxt = nxt, cur
            reqs = self._exchange(cur, nxt) if t < n - 1 else []
            for qh, kh in self._pairs(t):
                work[qh][:, :, :, 1:].copy_(cur[kh])
                o_j, lse_j = attention_fwd(work[qh][:, :, :, 0], work[qh][:, :, :, 1], work[qh][:, :, :, 2], scale, False)
                attn_merge_(o_acc[qh], lse_a[qh], lse_b[qh], o_j.contiguous(), lse_j.contiguous(), False)
                lse_a[qh], lse_b[qh] = lse_b[qh], lse_a[qh]
                self.block_products += 0.25
        self._wait(reqs)
        o = torch.cat((o_acc[0], o_acc[1]), 1).to(q.dtype)
        lse = torch.cat((lse_a[0], lse_a[1]), 2)
        # the output goes back to the graph's (contiguous) layout; the log-sum-exp is only ever read by the matching backward,
        # which works in the zig-zag layout: it stays there
        return self._permute(o, False), lse

    def _backward_zigzag(self, do, q, k, v, o, lse, scale, dqkv_out):
        B, L, H, D = q.shape
        n, l, dev = self.n, L // 2, q.device
        qkv = self._permute(torch.stack((q, k, v), 3), True)
        go = self._permute(torch.stack((do, o), 3), True)                      # [B,L,H,2,D]
        qz, kz, vz = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
        doz, oz = go[:, :, :, 0].contiguous(), go[:, :, :, 1].contiguous()
        lse = lse.contiguous()
        part = torch.empty(B, L, H, 3, D, dtype=q.dtype, device=dev)
        attention_bwd(doz, qz, kz, vz, oz, lse, scale, True, dqkv_out=part)    # local step
        self.block_products += 0.5
        ph = self._halves(part)                                                # [2,B,l,H,3,D]
        dq_acc = ph[:, :, :, :, 0].float().contiguous()                        # [2,B,l,H,D]
        acc = ph[:, :, :, :, 1:].float().contiguous()                          # [2,B,l,H,2,D]: dK / dV of the block `cur` holds
        acc_in = torch.empty_like(acc)
        cur = self._halves(torch.stack((kz, vz), 3))
        nxt = torch.empty_like(cur)
        doh, oh = self._halves(doz), self._halves(oz)
        lh = (lse[:, :, :l].contiguous(), lse[:, :, l:].contiguous())
        work = [torch.empty(B, l, H, 3, D, dtype=q.dtype, device=dev) for _ in range(2)]
        for h in range(2):
            work[h][:, :, :, 0].copy_(qz[:, h * l:(h + 1) * l])
        small = torch.empty(B, l, H, 3, D, dtype=q.dtype, device=dev)
        reqs = self._exchange(cur, nxt)
        for t in range(1, n + 1):
            self._wait(reqs)
            self._wait(self._exchange(acc, acc_in))                            # the accumulator follows its block
            acc, acc_in = acc_in, acc
            if t == n:
                break                                                          # acc now holds this rank's own dK / dV
            cur, nxt = nxt, cur
            reqs = self._exchange(cur, nxt) if t < n - 1 else []
            for qh, kh in self._pairs(t):
                work[qh][:, :, :, 1:].copy_(cur[kh])
                attention_bwd(doh[qh], work[qh][:, :, :, 0], work[qh][:, :, :, 1], work[qh][:, :, :, 2], oh[qh], lh[qh], scale, False,
                              dqkv_out=small)
                attn_ring_accum_(dq_acc[qh], acc[kh], small)
                self.block_products += 0.25
        gz = torch.empty(B, L, H, 3, D, dtype=q.dtype, device=dev)
        for h in range(2):
            gz[:, h * l:(h + 1) * l, :, 0].copy_(dq_acc[h])
            gz[:, h * l:(h + 1) * l, :, 1:].copy_(acc[h])
        g = self._permute(gz, False)
        dqkv_out.copy_(g)
        return dqkv_out[:, :, :, 0], dqkv_out[:, :, :, 1], dqkv_out[:, :, :, 2]
