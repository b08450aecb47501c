"""GPT-MoE (reference examples/gpt_moe: GShard-style top-2 gating, E experts, einsum dispatch/combine) in planner
IR.  Expert parallelism is *emergent*: the expert FFN einsums carry a leading expert dim E, so the planner's batch-
split proposal on E is expert parallel, and the dispatch/combine einsums change the split dim G <-> E, which makes
the planner insert all-to-all (SURVEY §2.G "EP")."""
from __future__ import annotations

import math
from dataclasses import dataclass

from ..frontend.builder import GraphBuilder, build_training_step
from ..ir import Graph, Value


@dataclass
class MoEConfig:
    n_layer: int = 8
    hidden: int = 768
    ffn: int = 6144
    n_head: int = 16
    experts: int = 8
    capacity: int = 256
    groups: int = 8
    seq: int = 1024
    batch: int = 8
    vocab: int = 50257
    name: str = "gpt-moe-8e"

    @property
    def padded_vocab(self) -> int:
        return (self.vocab + 127) // 128 * 128


def moe_ffn(b: GraphBuilder, x: Value, groups: int, experts: int, capacity: int, hidden: int, name: str = "moe") -> Value:
    """x: [G, S, M] (tokens grouped).  Returns [G, S, M].  Gating is top-1 'switch' style computed with softmax; the
    dispatch mask [G,S,E,C] is produced by the runtime gating op; dispatch / combine are einsums so the planner sees
    the G <-> E re-distribution."""
    G, S_, M = b.t(x).shape
    E, C, H = experts, capacity, hidden
    nrm = lambda s: {"kind": "normal", "mean": 0.0, "std": s}
    with b.scope(name):
        wg = b.parameter("gate/w", (E, M), nrm(0.02))
        logits = b.linear(x, wg, name="gate")                                # [G,S,E]
        gates = b.softmax(b.cast(logits, "f32"), -1, name="gate_softmax")
        combine = b.moe_dispatch_mask(gates, C, top_k=2, dtype=b.t(x).dtype)   # combine weights [G,S,E,C] (0 where dropped)
        dispatched = b.einsum("GSEC,GSM->EGCM", combine, x, name="dispatch")  # tokens to experts: G-major -> E-major
        wi = b.parameter("expert_fc/wi", (E, M, H), nrm(0.02))
        wo = b.parameter("expert_fc/wo", (E, H, M), nrm(0.02 / math.sqrt(2)))
        h = b.einsum("EGCM,EMH->EGCH", dispatched, wi, name="expert_fc1")
        h = b.gelu(h, name="expert_gelu")
        y = b.einsum("EGCH,EHM->EGCM", h, wo, name="expert_fc2")
        out = b.einsum("GSEC,EGCM->GSM", combine, y, name="combine")         # back to G-major
    return out


def build_moe_ffn_graph(groups=8, tokens_per_group=64, model=64, hidden=256, experts=8, capacity=16) -> Graph:
    """One MoE FFN layer with an MSE loss — the smallest graph that exercises EP + all-to-all planning."""
    b = GraphBuilder("moe_ffn")
    x = b.input("x", (groups, tokens_per_group, model), "bf16")
    t = b.input("t", (groups, tokens_per_group, model), "bf16")
    y = moe_ffn(b, x, groups, experts, capacity, hidden)
    d = b.cast(b.sub(y, t), "f32")
    loss = b.reduce_mean(b.mul(d, d), [0, 1, 2], name="loss")
    return build_training_step(b, loss, "adamw", lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0)


def build_gpt_moe_graph(cfg: MoEConfig, batch: int | None = None, optimizer: str = "adamw", **opt_hp) -> Graph:
    """`opt_hp`: further optimizer settings of pretrain_moe.json, e.g. clip_norm="global", clip_norm_value=1.0, schedule={...}."""
    """Full GPT-MoE: every other layer's MLP is an MoE FFN (pretrain_moe.json: 8 layers, hidden 768, 16 heads x 48,
    8 experts, capacity 256, 8 local groups, seq 1024)."""
    B = batch or cfg.batch
    S_, C, H, Vp = cfg.seq, cfg.hidden, cfg.n_head, cfg.padded_vocab
    b = GraphBuilder(cfg.name)
    tokens = b.input("tokens", (B, S_), "i32")
    labels = b.input("labels", (B, S_), "i32")
    nrm = lambda s: {"kind": "normal", "mean": 0.0, "std": s}
    const = lambda v: {"kind": "constant", "value": v}
    with b.scope("model"):
        wte = b.parameter("wte", (Vp, C), nrm(0.02))
        wpe = b.parameter("wpe", (S_, C), nrm(0.01))
        x = b.embedding(tokens, wte, wpe)
        for l in range(cfg.n_layer):
            with b.scope(f"h{l}"):
                g1 = b.parameter("ln_1/g", (C,), const(1.0)); b1 = b.parameter("ln_1/b", (C,), const(0.0))
                h = b.layernorm(x, g1, b1, name="ln_1")
                wq = b.parameter("attn/c_attn/w", (3 * C, C), nrm(0.02)); bq = b.parameter("attn/c_attn/b", (3 * C,), const(0.0))
                a = b.attention(b.linear(h, wq, bq, name="attn/c_attn"), heads=H, causal=True, name="attn/core")
                wo = b.parameter("attn/c_proj/w", (C, C), nrm(0.02)); bo = b.parameter("attn/c_proj/b", (C,), const(0.0))
                x = b.linear(a, wo, bo, residual=x, name="attn/c_proj")
                g2 = b.parameter("ln_2/g", (C,), const(1.0)); b2 = b.parameter("ln_2/b", (C,), const(0.0))
                h = b.layernorm(x, g2, b2, name="ln_2")
                if l % 2 == 1:
                    tokens_total = B * S_
                    G = cfg.groups
                    hg = b.reshape(h, (G, tokens_total // G, C), name="to_groups")
                    m = moe_ffn(b, hg, G, cfg.experts, cfg.capacity, cfg.ffn, name="moe")
                    x = b.add(x, b.reshape(m, (B, S_, C), name="from_groups"), name="moe_res")
                else:
                    wf = b.parameter("mlp/c_fc/w", (cfg.ffn, C), nrm(0.02)); bf_ = b.parameter("mlp/c_fc/b", (cfg.ffn,), const(0.0))
                    f = b.gelu(b.linear(h, wf, bf_, name="mlp/c_fc"))
                    wp = b.parameter("mlp/c_proj/w", (C, cfg.ffn), nrm(0.02)); bp = b.parameter("mlp/c_proj/b", (C,), const(0.0))
                    x = b.linear(f, wp, bp, residual=x, name="mlp/c_proj")
        gf = b.parameter("ln_f/g", (C,), const(1.0)); bf2 = b.parameter("ln_f/b", (C,), const(0.0))
        logits = b.linear(b.layernorm(x, gf, bf2, name="ln_f"), b.parameter("output", (Vp, C), nrm(0.02)), name="lm_head")
        loss = b.softmax_xent(logits, labels, vocab=cfg.vocab, name="loss")
    # (pretrain_moe.json "optimizer": adamw | adafactor | lamb | sm3 -- frontend/builder.py OPTIMIZERS)
    return build_training_step(b, loss, optimizer, **{**dict(lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01), **opt_hp})
