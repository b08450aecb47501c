#!/usr/bin/env python
"""GPT-MoE (reference examples/gpt_moe/run_moe.sh: 8 layers, hidden 768, 8 experts, top-2 gating, capacity 256)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tepdist_b200.api import Trainer  # noqa: E402
from tepdist_b200.models.gpt_moe import MoEConfig, build_gpt_moe_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--strategy", default="auto")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "adafactor", "lamb", "sm3"],
                    help='reference: "optimizer" in examples/gpt_moe/pretrain_moe.json')
    ap.add_argument("--clip-norm", default=None, choices=["global", "local"], help='gradient clipping (reference gpt_moe config: "clip_norm")')
    ap.add_argument("--clip-norm-value", type=float, default=1.0)
    ap.add_argument("--data", default=None, help="comma-separated token files (tepdist_b200.data.write_token_file) or 'synthetic'; default: one fixed random batch")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly (default: whole step replayed from a CUDA graph)")
    a = ap.parse_args()
    clip = {"clip_norm": a.clip_norm, "clip_norm_value": a.clip_norm_value} if a.clip_norm else {}
    cfg = MoEConfig(batch=a.batch)
    if a.tiny:
        cfg = MoEConfig(n_layer=2, hidden=128, ffn=256, n_head=2, experts=4, capacity=64, groups=4, seq=128, batch=a.batch, vocab=1000)
    tr = Trainer(build_gpt_moe_graph(cfg, optimizer=a.optimizer, **clip), strategy=a.strategy, use_cuda_graph=not a.no_graph)
    if a.data:      # native prefetching loader: every rank draws the same global batch, the trainer takes what its plan assigns to it
        from tepdist_b200.data import TokenLoader
        src = {"synthetic_vocab": cfg.vocab} if a.data == "synthetic" else {"files": a.data.split(",")}
        batches = iter(TokenLoader(batch=cfg.batch, n_ctx=cfg.seq, **src))
    else:
        gen = torch.Generator().manual_seed(0)
        tok = torch.randint(0, cfg.vocab, (cfg.batch, cfg.seq), generator=gen, dtype=torch.int32)
        fixed = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
        batches = iter(lambda: fixed, None)
    t0 = time.time()
    t_steady = None
    for i in range(a.steps):
        if i == 3:      # steps 0-2: planning artefacts, lazy allocations, kernel attribute setup
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            t_steady = time.time()
        loss = tr.step(next(batches))
        if tr.rank == 0:
            print(f"step {i} loss {loss:.4f}")
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if tr.rank == 0:
        msg = f"{a.steps} steps in {time.time() - t0:.2f}s; plan: {tr.plan_info.get('parallelism', 'single')}"
        if t_steady is not None and a.steps > 3:
            dt = (time.time() - t_steady) / (a.steps - 3)
            msg += f"; steady state {dt * 1e3:.1f} ms/step = {cfg.batch * cfg.seq / dt:.0f} tokens/s (global batch {cfg.batch} x {cfg.seq})"
        msg += f"; collectives: {tr.plan_info.get('collectives')}"
        print(msg)
    os._exit(0)


if __name__ == "__main__":
    main()
