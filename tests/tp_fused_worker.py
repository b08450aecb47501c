"""2+ GPU worker: tensor-parallel fused GEMM kernels over peer memory vs plain references, plus timing vs NCCL."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ev_ms(fn, iters=20, warm=5, graph=True):
    """Device time per call, max over ranks.  The calls are replayed from a CUDA graph (20 per replay) so the number
    is the kernels' time, not the Python launch path's (every variant issues 3-4 launches per call)."""
    for _ in range(warm):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    g = None
    if graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(iters):
                        fn()
            torch.cuda.current_stream().wait_stream(s)
            g.replay()
        except Exception as e:  # noqa: BLE001
            print("graph capture failed, timing eager:", type(e).__name__, e, flush=True)
            g = None
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if g is not None else 1
    e0.record()
    for _ in range(reps):
        if g is not None:
            g.replay()
        else:
            for _ in range(iters):
                fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (iters * reps)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    out = sys.argv[1]
    from tepdist_b200 import ops
    from tepdist_b200.api import init_distributed
    from tepdist_b200.parallel.symm import (AllGatherGemm, GemmReduceScatter, SymmBarrier, SymmetricBuffer, all_gather_gemm,
                                            gemm_reduce_scatter)
    ctx = init_distributed()
    rank, n = ctx["rank"], ctx["world"]
    dev = torch.device("cuda", ctx["local_rank"])
    res = {}
    M, N, K = 4096, 1024, 4096          # row-parallel c_proj of a 1024-wide MLP: K sharded
    Kl = K // n
    torch.manual_seed(0)
    X = torch.randn(M, K, device=dev, dtype=torch.bfloat16) * 0.5
    W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    xs, ws = X[:, rank * Kl:(rank + 1) * Kl].contiguous(), W[:, rank * Kl:(rank + 1) * Kl].contiguous()
    bar = SymmBarrier()
    # ---- GEMM -> reduce-scatter
    acc = SymmetricBuffer((M // n) * N * 4)
    acc_t = acc.tensor(torch.float32, (M // n) * N)
    acc_t.zero_()
    bar()
    y = gemm_reduce_scatter(xs, ws, acc, N)
    bar()
    torch.cuda.synchronize()
    ref = (X.float() @ W.float().t())[rank * (M // n):(rank + 1) * (M // n)]
    res["gemm_rs_relerr"] = float((y - ref).norm() / ref.norm())

    def fused():
        acc_t.zero_(); bar(); gemm_reduce_scatter(xs, ws, acc, N); bar()

    part = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    shard = torch.empty(M // n, N, device=dev, dtype=torch.bfloat16)

    def nccl():
        torch.matmul(xs, ws.t(), out=part)
        dist.reduce_scatter_tensor(shard, part)

    res["gemm_rs_fused_ms"] = ev_ms(fused)
    res["gemm_rs_nccl_ms"] = ev_ms(nccl)
    # slot variant: bf16 partials pushed with plain stores, summed on the owner together with bias + residual
    grs = GemmReduceScatter(M, N, barrier=bar)
    bias = torch.randn(N, device=dev, dtype=torch.float32)
    resid = torch.randn(M // n, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):                       # exercise both staging buffers
        y3 = grs(xs, ws, bias=bias, residual=resid)
    torch.cuda.synchronize()
    ref3 = ref + bias + resid.float()
    res["gemm_rs_slots_relerr"] = float((y3.float() - ref3).norm() / ref3.norm())
    res["gemm_rs_slots_ms"] = ev_ms(lambda: grs(xs, ws, bias=bias, residual=resid))

    def nccl_full():
        torch.matmul(xs, ws.t(), out=part)
        dist.reduce_scatter_tensor(shard, part)
        torch.add(shard, resid, out=shard)
        shard.add_(bias.to(torch.bfloat16))

    res["gemm_rs_nccl_bias_res_ms"] = ev_ms(nccl_full)
    # ---- all-gather -> GEMM (column-parallel c_fc with row-sharded activations)
    M2, K2, N2 = 4096, 1024, 4096 // n
    A = torch.randn(M2, K2, device=dev, dtype=torch.bfloat16) * 0.5
    Wc = torch.randn(N2, K2, device=dev, dtype=torch.bfloat16) * 0.05
    sh = SymmetricBuffer((M2 // n) * K2 * 2)
    sh.tensor(torch.bfloat16, (M2 // n) * K2).view(M2 // n, K2).copy_(A[rank * (M2 // n):(rank + 1) * (M2 // n)])
    bar()
    d = all_gather_gemm(sh, M2 // n, K2, Wc)
    bar()
    torch.cuda.synchronize()
    ref2 = A.float() @ Wc.float().t()
    res["ag_gemm_relerr"] = float((d.float() - ref2).norm() / ref2.norm())
    full = torch.empty(M2, K2, device=dev, dtype=torch.bfloat16)
    mine = A[rank * (M2 // n):(rank + 1) * (M2 // n)].contiguous()

    def nccl2():
        dist.all_gather_into_tensor(full, mine)
        torch.matmul(full, Wc.t())

    res["ag_gemm_fused_ms"] = ev_ms(lambda: (bar(), all_gather_gemm(sh, M2 // n, K2, Wc)))
    res["ag_gemm_nccl_ms"] = ev_ms(nccl2)
    # staged variant: copy kernel + flag-gated persistent GEMM
    agg = AllGatherGemm(M2 // n, K2, barrier=bar)
    for it in range(3):
        agg.input().copy_(A[rank * (M2 // n):(rank + 1) * (M2 // n)] * (it + 1))
        d4 = agg(Wc)
    torch.cuda.synchronize()
    res["ag_gemm_staged_relerr"] = float((d4.float() - 3 * ref2).norm() / (3 * ref2).norm())
    res["ag_gemm_staged_ms"] = ev_ms(lambda: agg(Wc))
    res["world"] = n
    if rank == 0:
        json.dump(res, open(out, "w"))
        print("TPFUSED", json.dumps(res))
    dist.barrier(); torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
