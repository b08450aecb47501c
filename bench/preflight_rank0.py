#!/usr/bin/env python
"""Pre-flight of a multi-GPU plan WITHOUT GPUs: builds rank 0's executor for the N-way plan of the benchmark model on the CPU
(a stand-in mesh, no process group) and runs steps with the communication replaced by its shape-preserving stand-ins (the same
`dry` mode bench.py uses to measure exposed communication).  Every Python-level code path the real N-GPU run takes -- planning,
sharded-optimizer detection, bucket layout, slot allocation, the per-bucket update loop, state_dict -- executes; only the
CUDA kernels and the real collectives do not.  Costs CPU time and ~6 GB for GPT-2 345M, no GPU minutes:

    python bench/preflight_rank0.py --model 345M --gpus 8 --steps 1

(The printed loss is this rank's partial loss: local tokens over the GLOBAL token count, i.e. ~1/N of the usual value.)
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph  # noqa: E402
from tepdist_b200.parallel import plan_spmd  # noqa: E402
from tepdist_b200.parallel.collectives import CollectiveRunner  # noqa: E402
from tepdist_b200.runtime.executor import Executor  # noqa: E402


class StandInMesh:
    def __init__(self, world):
        self.world, self.rank = world, 0

    def group(self, level):
        return None

    def index_in_group(self, level):
        return 0

    def coords(self, d=None):
        return [0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="345M", choices=list(CONFIGS))
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4, help="sequences per GPU")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--strategy", default="auto")
    a = ap.parse_args()
    cfg = CONFIGS[a.model]
    g = build_gpt2_graph(cfg, batch=a.batch * a.gpus)
    t = time.time()
    g2, info = plan_spmd(g, a.gpus, a.strategy)
    print(f"plan: {info['collectives']}  ({time.time() - t:.2f} s)")
    col = CollectiveRunner(StandInMesh(a.gpus))
    t = time.time()
    ex = Executor(g2, torch.device("cpu"), seed=0, collective=col, coords={0: 0}, comm_mode="nccl", use_cuda_graph=False)
    st, fz = ex.store, ex.flat_zero
    print(f"executor: {time.time() - t:.1f} s  sharded optimizer: {fz is not None}  flat m/v: {st.m is not None}  "
          f"buckets: {len(fz['buckets']) if fz else None}  regular / replicated / all updates: "
          f"{len(fz['regular_apply']) if fz else None} / {len(fz.get('replicated_apply') or {}) if fz else None} / {len(ex.apply_nodes)}")
    ex.dry_comm = col.dry = True
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (a.batch, cfg.n_ctx), dtype=torch.int32)
    for i in range(a.steps):
        t = time.time()
        loss = float(ex.step({"tokens": tok, "labels": torch.roll(tok, -1, 1)})[0])
        print(f"step {i}: partial loss {loss:.4f}  ({time.time() - t:.1f} s)")
    print(f"state_dict: {len(st.state_dict())} entries")


if __name__ == "__main__":
    main()
