"""Smoke-test models mirroring the reference's examples/smoke_testing/{simple,attention,conv}.py."""
from __future__ import annotations

from ..frontend.builder import GraphBuilder, build_training_step
from ..ir import Graph


def build_mlp_graph(batch: int = 8, d_in: int = 16, d_hidden: int = 32, d_out: int = 4, lr: float = 0.5,
                    annotate: bool = False, num: int = 2) -> Graph:
    """2-layer MLP -> softmax -> MSE, SGD (simple.py: matmul -> softmax -> sum, SGD, optional xla_sharding.split)."""
    b = GraphBuilder("smoke_mlp", compute_dtype="f32")
    x = b.input("x", (batch, d_in), "f32")
    t = b.input("t", (batch, d_out), "f32")
    w1 = b.parameter("w1", (d_in, d_hidden), {"kind": "normal", "std": 0.3})
    w2 = b.parameter("w2", (d_hidden, d_out), {"kind": "normal", "std": 0.3})
    if annotate:
        b.annotate_split(w1, 1, num)
    h = b.tanh(b.matmul(x, w1, name="fc1"))
    y = b.softmax(b.matmul(h, w2, name="fc2"))
    d = b.sub(y, t)
    loss = b.reduce_mean(b.mul(d, d), [0, 1], name="loss")
    return build_training_step(b, loss, "sgd", lr=lr)


def build_attention_graph(batch: int = 2, seq: int = 128, hidden: int = 128, heads: int = 2) -> Graph:
    """Single attention block (attention.py)."""
    b = GraphBuilder("smoke_attention")
    x = b.input("x", (batch, seq, hidden), "bf16")
    t = b.input("t", (batch, seq, hidden), "bf16")
    wq = b.parameter("c_attn/w", (3 * hidden, hidden), {"kind": "normal", "std": 0.05})
    wo = b.parameter("c_proj/w", (hidden, hidden), {"kind": "normal", "std": 0.05})
    a = b.attention(b.linear(x, wq, name="c_attn"), heads=heads, causal=True)
    y = b.linear(a, wo, name="c_proj")
    d = b.cast(b.sub(y, t), "f32")
    loss = b.reduce_mean(b.mul(d, d), [0, 1, 2], name="loss")
    return build_training_step(b, loss, "sgd", lr=0.1)


def build_conv_graph(batch: int = 4, image: int = 16, channels: int = 8, classes: int = 10) -> Graph:
    """Small conv net (conv.py): conv -> bn -> relu -> conv -> gap -> fc -> mse."""
    b = GraphBuilder("smoke_conv", compute_dtype="f32")
    x = b.input("x", (batch, 3, image, image), "f32")
    t = b.input("t", (batch, classes), "f32")
    w1 = b.parameter("conv1/w", (channels, 3, 3, 3), {"kind": "normal", "std": 0.2})
    g1 = b.parameter("bn1/g", (channels,), {"kind": "constant", "value": 1.0})
    b1 = b.parameter("bn1/b", (channels,), {"kind": "constant", "value": 0.0})
    w2 = b.parameter("conv2/w", (2 * channels, channels, 3, 3), {"kind": "normal", "std": 0.1})
    wf = b.parameter("fc/w", (2 * channels, classes), {"kind": "normal", "std": 0.2})
    h = b.relu(b.batchnorm(b.conv2d(x, w1, 1, 1), g1, b1))
    h = b.relu(b.conv2d(h, w2, 2, 1))
    y = b.matmul(b.global_avgpool(h), wf, name="fc")
    d = b.sub(y, t)
    loss = b.reduce_mean(b.mul(d, d), [0, 1], name="loss")
    return build_training_step(b, loss, "sgd", lr=0.05)
