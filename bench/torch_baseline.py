"""Reference-semantics baseline for bench.py (`--impl library`): the SAME GPT-2 training step executed the way the
reference executes it on a box, with none of this repository's kernels, planner or runtime on the path.

What "reference semantics" means ("What the comparison target therefore is"):
  * library GEMMs (cuBLAS through torch), library attention (SDPA / cuDNN-flash), library elementwise kernels;
  * data parallelism with one in-stream ncclAllReduce PER GRADIENT TENSOR after the backward pass, no bucketing,
    no overlap, no reduce-scatter, every rank runs the full optimizer (dapple_all_reduce_thunk.cc:136-159);
  * same numerics policy as our arm: bf16 compute, fp32 master weights + AdamW moments, same architecture
    (untied lm_head, vocab padded to 128), same per-GPU batch and sequence length, synthetic tokens.
The whole step is captured into a CUDA graph when possible (the reference's XLA executable has no per-op Python
overhead, so an eager-mode Python loop would be an unfairly weak baseline); falls back to eager if capture fails.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
import torch.nn.functional as F


class LibraryGPT2:
    def __init__(self, cfg, device, seed: int = 0):
        C, L, S, Vp = cfg.n_embd, cfg.n_layer, cfg.n_ctx, cfg.padded_vocab
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.cfg, self.dev = cfg, device
        self.params = {}

        def nrm(name, shape, std):
            self.params[name] = (torch.randn(shape, generator=g) * std).to(device).requires_grad_(True)

        def const(name, shape, v):
            self.params[name] = torch.full(shape, v, device=device).requires_grad_(True)

        nrm("wte", (Vp, C), 0.02); nrm("wpe", (S, C), 0.01)
        for l in range(L):
            p = f"h{l}/"
            const(p + "ln1g", (C,), 1.0); const(p + "ln1b", (C,), 0.0)
            nrm(p + "qkv_w", (3 * C, C), 0.02); const(p + "qkv_b", (3 * C,), 0.0)
            nrm(p + "o_w", (C, C), 0.02 / math.sqrt(2 * L)); const(p + "o_b", (C,), 0.0)
            const(p + "ln2g", (C,), 1.0); const(p + "ln2b", (C,), 0.0)
            nrm(p + "fc_w", (4 * C, C), 0.02); const(p + "fc_b", (4 * C,), 0.0)
            nrm(p + "pr_w", (C, 4 * C), 0.02 / math.sqrt(2 * L)); const(p + "pr_b", (C,), 0.0)
        const("lnfg", (C,), 1.0); const("lnfb", (C,), 0.0)
        nrm("out_w", (Vp, C), 0.02)
        decay = [v for k, v in self.params.items() if v.dim() >= 2]
        no_decay = [v for k, v in self.params.items() if v.dim() < 2]
        self.opt = torch.optim.AdamW([{"params": decay, "weight_decay": cfg.weight_decay},
                                      {"params": no_decay, "weight_decay": 0.0}], lr=cfg.lr, betas=(0.9, 0.999), eps=1e-8,
                                     fused=True, capturable=True)

    def loss(self, tokens, labels):
        cfg, P = self.cfg, self.params
        B, S = tokens.shape
        C, H = cfg.n_embd, cfg.n_head
        with torch.autocast(self.dev.type, dtype=torch.bfloat16):
            x = (F.embedding(tokens, P["wte"]) + P["wpe"][:S]).to(torch.bfloat16)
            for l in range(cfg.n_layer):
                p = f"h{l}/"
                h = F.layer_norm(x, (C,), P[p + "ln1g"], P[p + "ln1b"])
                qkv = F.linear(h, P[p + "qkv_w"], P[p + "qkv_b"]).view(B, S, 3, H, C // H)
                q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
                a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, C)
                x = x + F.linear(a, P[p + "o_w"], P[p + "o_b"])
                h = F.layer_norm(x, (C,), P[p + "ln2g"], P[p + "ln2b"])
                f = F.gelu(F.linear(h, P[p + "fc_w"], P[p + "fc_b"]), approximate="tanh")
                x = x + F.linear(f, P[p + "pr_w"], P[p + "pr_b"])
            h = F.layer_norm(x, (C,), P["lnfg"], P["lnfb"])
            logits = F.linear(h, P["out_w"])[..., :cfg.n_vocab]
        return F.cross_entropy(logits.float().view(-1, cfg.n_vocab), labels.view(-1).long())


class LibraryTrainer:
    """step(tokens, labels) -> loss tensor; whole step (fwd, bwd, per-tensor all-reduce, AdamW) in one CUDA graph."""

    def __init__(self, cfg, device, world: int, use_graph: bool = True):
        self.model = LibraryGPT2(cfg, device)
        self.world, self.dev = world, device
        self.use_graph = use_graph
        self.graph = None
        self.launches_per_step = None
        self.static_tok = self.static_lab = self.static_loss = None

    def _step_body(self, tok, lab):
        m = self.model
        loss = m.loss(tok.long(), lab)
        loss.backward()
        if self.world > 1:
            for p in m.params.values():           # one in-stream all-reduce per gradient tensor (reference K1)
                dist.all_reduce(p.grad)
                p.grad.div_(self.world)
        m.opt.step()
        m.opt.zero_grad(set_to_none=False)
        return loss.detach()

    def step(self, tok, lab):
        if not self.use_graph:
            return self._step_body(tok, lab)
        if self.graph is None:
            self.static_tok, self.static_lab = tok.clone(), lab.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._step_body(self.static_tok, self.static_lab)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.static_loss = self._step_body(self.static_tok, self.static_lab)
                self.graph = g
            except Exception as e:  # noqa: BLE001
                print(f"[library baseline] CUDA-graph capture failed ({type(e).__name__}: {e}); running eager", flush=True)
                self.use_graph = False
                torch.cuda.synchronize()
                return self._step_body(tok, lab)
        self.static_tok.copy_(tok, non_blocking=True)
        self.static_lab.copy_(lab, non_blocking=True)
        self.graph.replay()
        return self.static_loss
