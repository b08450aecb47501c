"""Per-op device time of one GPT-2 training step (Executor.profile: CUDA events around every node of an eager step).
usage: python bench/profile_step.py [--model 345M] [--batch 4] [--out gpurun_out/step_profile.json]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph  # noqa: E402
from tepdist_b200.runtime.executor import Executor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="345M")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--out", default="gpurun_out/step_profile.json")
a = ap.parse_args()
cfg = CONFIGS[a.model]
ex = Executor(build_gpt2_graph(cfg, batch=a.batch), torch.device("cuda", 0), use_cuda_graph=False)
tok = torch.randint(0, cfg.n_vocab, (a.batch, cfg.n_ctx), dtype=torch.int32, device="cuda")
feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
prof = ex.profile(feeds, warmup=3, chrome_trace=a.out.replace(".json", "_trace.json"))
tot = sum(prof["by_op"].values())
print(f"sum of node times {tot:.2f} ms (eager wall {prof['total_ms']:.2f} ms incl. host launch gaps)")
cnt = {}
for r in prof["nodes"]:
    cnt[r["op"]] = cnt.get(r["op"], 0) + 1
for op, ms in prof["by_op"].items():
    print(f"{op:18s} {cnt.get(op, 0):4d} x  {ms:8.3f} ms  {100 * ms / tot:5.1f} %   avg {1e3 * ms / max(1, cnt.get(op, 0)):7.1f} us")
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump({"by_op_ms": prof["by_op"], "counts": cnt, "sum_ms": tot}, open(a.out, "w"), indent=1)
