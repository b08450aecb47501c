"""Tiny driver for ncu: launches the flagship GEMM shapes + attention + fused AdamW a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepdist_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
if which == "gemm":
    M, N, K = 4096, 4096, 1024
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    W = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(12):
        ops.gemm(A, W, block_n=256)
elif which == "gemm2":
    M, N, K = 4096, 4096, 1024
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    W = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(12):
        ops.gemm2(A, W)
elif which == "attn":
    qkv = torch.randn(4, 1024, 16, 3, 64, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
    for _ in range(6):
        o, lse = ops.attention_fwd(q, k, v)
        ops.attention_bwd(torch.randn_like(o), q, k, v, o, lse)
torch.cuda.synchronize()
