#!/usr/bin/env python
"""Bring-your-own torch.nn.Module: a small GPT written in plain PyTorch (nn.Embedding / nn.LayerNorm / nn.Linear /
F.scaled_dot_product_attention / F.gelu) is traced with torch.fx into ONE training-step graph (forward + backward + AdamW),
planned automatically and executed by the same runtime as the built-in models.  Counterpart of the reference's "unmodified
TensorFlow model + session.run(train_op)" workflow (README.md:101-143)."""
import argparse
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tepdist_b200.api import Trainer  # noqa: E402
from tepdist_b200.frontend.trace import trace  # noqa: E402


class Block(nn.Module):
    def __init__(self, C, H, B, S):
        super().__init__()
        self.C, self.H, self.B, self.S = C, H, B, S
        self.ln1, self.ln2 = nn.LayerNorm(C), nn.LayerNorm(C)
        self.qkv, self.proj = nn.Linear(C, 3 * C), nn.Linear(C, C)
        self.fc, self.out = nn.Linear(C, 4 * C), nn.Linear(4 * C, C)

    def forward(self, x):
        q, k, v = self.qkv(self.ln1(x)).chunk(3, dim=-1)
        q, k, v = (t.view(self.B, self.S, self.H, self.C // self.H).transpose(1, 2) for t in (q, k, v))
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(self.B, self.S, self.C)
        x = x + self.proj(a)
        return x + self.out(F.gelu(self.fc(self.ln2(x)), approximate="tanh"))


class TinyGPT(nn.Module):
    def __init__(self, V=1024, C=256, H=4, L=2, B=4, S=128):
        super().__init__()
        self.emb = nn.Embedding(V, C)
        self.blocks = nn.Sequential(*[Block(C, H, B, S) for _ in range(L)])
        self.ln_f = nn.LayerNorm(C)
        self.head = nn.Linear(C, V, bias=False)

    def forward(self, tokens):
        return self.head(self.ln_f(self.blocks(self.emb(tokens))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--strategy", default="auto")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, S, V = 4 * world, 128, 1024
    net = TinyGPT(V=V, B=B, S=S)
    tok = torch.randint(0, V, (B, S))
    dev_cuda = torch.cuda.is_available()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.01)     # the user's own optimizer object, taken as it is
    tr = trace(net, {"tokens": tok}, loss="cross_entropy", label_example=tok.int(), optimizer=opt,
               compute_dtype="bf16" if dev_cuda else "f32")
    trainer = Trainer(tr.graph, strategy=a.strategy, use_cuda_graph=False)
    tr.load_state_dict_into(trainer.exec, net.state_dict())       # start from the module's own weights
    feeds = {"tokens": tok.int(), "labels": torch.roll(tok, -1, 1).int()}
    for i in range(a.steps):
        loss = trainer.step(feeds)
        if trainer.rank == 0:
            print(f"step {i} loss {loss:.4f}  plan: {trainer.plan_info.get('parallelism', 'single')}")
    os._exit(0)


if __name__ == "__main__":
    main()
