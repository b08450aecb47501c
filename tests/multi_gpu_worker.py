"""Multi-GPU worker (torchrun): planner-driven DP/ZeRO-1 with the fused peer-memory optimizer kernel vs the NCCL
reference-semantics path vs a single-GPU run of the same global batch."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    out = sys.argv[1]
    from tepdist_b200 import ops
    from tepdist_b200.api import Trainer
    from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph
    from tepdist_b200.runtime.executor import Executor
    cfg = CONFIGS["tiny"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = 4 * world
    g = build_gpt2_graph(cfg, batch=B)
    res = {}
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (B, cfg.n_ctx), dtype=torch.int32)
    lab = torch.roll(tok, -1, 1)
    for mode in ("nccl", "fused"):
        tr = Trainer(g, strategy="auto", use_cuda_graph=False, comm_mode=mode)
        n0 = ops.launch_count()
        res[mode] = [tr.step({"tokens": tok, "labels": lab}) for _ in range(5)]
        res[mode + "_launches"] = ops.launch_count() - n0
        res["parallelism"] = tr.plan_info.get("parallelism")
        res[mode + "_fused_active"] = bool(getattr(tr.exec, "flat_zero", None) and tr.exec.flat_zero.get("fused") is not None)
    if dist.get_rank() == 0:
        ref = Executor(g, torch.device("cuda", 0), seed=0)
        res["single"] = [float(ref.step({"tokens": tok.cuda(), "labels": lab.cuda()})[0]) for _ in range(5)]
        json.dump(res, open(out, "w"))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
