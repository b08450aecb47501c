"""In-situ kernel breakdown of the CUDA-graph training step: torch.profiler (CUPTI activity tracing) over a few graph
replays, aggregated per kernel name.  Shares of the real step (warm caches, back-to-back kernels) -- unlike the ncu launch
list, which serialises and runs every kernel cold.  usage: python bench/kineto_step.py [--model 345M] [--batch 4]"""
import argparse
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from tepdist_b200.api import Trainer  # noqa: E402
from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="345M")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/kineto_step.json")
a = ap.parse_args()
cfg = CONFIGS[a.model]
tr = Trainer(build_gpt2_graph(cfg, batch=a.batch))
tok = torch.randint(0, cfg.n_vocab, (a.batch, cfg.n_ctx), dtype=torch.int32, device=tr.device)
feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
for _ in range(6):
    tr.step_async(feeds)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        tr.step_async(feeds)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ka in prof.key_averages():
    # SELF device time: summing the inclusive field would count a kernel again under any parent range that got recorded
    us = getattr(ka, "self_device_time_total", None)
    if us is None:
        us = getattr(ka, "self_cuda_time_total", None)
    if us is None:
        us = getattr(ka, "device_time_total", 0.0)
    if not us:
        continue
    name = re.sub(r"^void ", "", ka.key)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = name.split("(")[0][:70]
    agg[name][0] += ka.count
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"kernel time per step {tot / a.steps / 1e3:.3f} ms over {sum(v[0] for v in agg.values()) // a.steps} kernels")
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for name, (c, us) in rows[:30]:
    print(f"{name:72s} {c // a.steps:4d} x {us / a.steps / 1e3:8.3f} ms {100 * us / tot:5.1f} %  avg {us / c:7.1f} us")
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump({"ms_per_step": tot / a.steps / 1e3, "kernels": {k: {"per_step": c / a.steps, "ms_per_step": us / a.steps / 1e3} for k, (c, us) in rows}},
          open(a.out, "w"), indent=1)
