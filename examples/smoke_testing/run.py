#!/usr/bin/env python
"""Smoke tests (reference examples/smoke_testing): mlp | attention | conv, 5 steps, optional sharding annotation / RPC."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tepdist_b200.api import Trainer  # noqa: E402
from tepdist_b200.models import smoke  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["mlp", "attention", "conv"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--strategy", default="auto")
    ap.add_argument("--annotate", action="store_true")
    ap.add_argument("--server", default=None, help="ip:port of a running tepdist_b200 server (client/server mode)")
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cpu") if a.cpu or not torch.cuda.is_available() else torch.device("cuda")
    torch.manual_seed(0)
    if a.which == "mlp":
        g = smoke.build_mlp_graph(batch=8, annotate=a.annotate)
        feeds = {"x": torch.randn(8, 16), "t": torch.rand(8, 4)}
    elif a.which == "attention":
        g = smoke.build_attention_graph()
        dt = torch.float32 if dev.type == "cpu" else torch.bfloat16
        feeds = {"x": torch.randn(2, 128, 128).to(dt), "t": torch.randn(2, 128, 128).to(dt)}
    else:
        g = smoke.build_conv_graph()
        feeds = {"x": torch.randn(4, 3, 16, 16), "t": torch.randn(4, 10)}
    if a.server:
        from tepdist_b200.rpc.client import Client
        cl = Client(a.server)
        print(cl.build_execution_plan(g, strategy=a.strategy)["plan_info"])
        for i in range(a.steps):
            print(f"step {i}: loss = {cl.execute_plan(feeds)['loss']:.6f}")
        return
    tr = Trainer(g, strategy=a.strategy, device=dev, use_cuda_graph=False)
    for i in range(a.steps):
        print(f"step {i}: loss = {tr.step(feeds):.6f}")


if __name__ == "__main__":
    main()
