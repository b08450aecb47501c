"""CPU tests: IR construction, autodiff vs torch.autograd, executor training loop."""
import pytest
import torch

from tepdist_b200.frontend.builder import GraphBuilder, build_training_step
from tepdist_b200.ir import Graph
from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph
from tepdist_b200.runtime.executor import Executor


def test_graph_roundtrip_json():
    g = build_gpt2_graph(CONFIGS["tiny"])
    g2 = Graph.from_json(g.to_json())
    assert len(g2.nodes) == len(g.nodes)
    assert [n.op for n in g2.nodes] == [n.op for n in g.nodes]
    assert g2.updates.keys() == g.updates.keys()
    g2.validate()


def test_op_groups_pair_forward_and_backward():
    g = build_gpt2_graph(CONFIGS["tiny"])
    fwd = {n.group: n for n in g.nodes if not n.backward and n.op == "linear"}
    for n in g.nodes:
        if n.op in ("linear_dgrad", "linear_wgrad"):
            assert n.backward and n.group in fwd
    # optimizer apply + slots take the variable's group
    for n in g.nodes:
        if n.op == "apply_adamw":
            assert n.group == g.nodes[n.inputs[0].node].group


def test_gpt2_tiny_loss_decreases_cpu():
    cfg = CONFIGS["tiny"]
    ex = Executor(build_gpt2_graph(cfg), torch.device("cpu"))
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    lab = torch.roll(tok, -1, 1)
    losses = [float(ex.step({"tokens": tok, "labels": lab})[0]) for _ in range(5)]
    assert losses[-1] < losses[0] - 0.1


def test_autodiff_matches_torch_autograd_mlp():
    """smoke_testing/simple.py shape: a[4,16] @ b[16,4] -> softmax -> sum, SGD."""
    b = GraphBuilder("simple", compute_dtype="f32")
    x = b.input("x", (8, 16), "f32")
    w1 = b.parameter("w1", (16, 32), {"kind": "normal", "std": 0.3})
    w2 = b.parameter("w2", (32, 4), {"kind": "normal", "std": 0.3})
    h = b.tanh(b.matmul(x, w1))
    y = b.softmax(b.matmul(h, w2))
    tgt = b.input("t", (8, 4), "f32")
    loss = b.reduce_mean(b.mul(b.sub(y, tgt), b.sub(y, tgt)), [0, 1])
    g = build_training_step(b, loss, "sgd", lr=0.5)
    ex = Executor(g, torch.device("cpu"))
    W1 = ex.store.master_view(w1.node).clone().requires_grad_(True)
    W2 = ex.store.master_view(w2.node).clone().requires_grad_(True)
    torch.manual_seed(1)
    X, T = torch.randn(8, 16), torch.rand(8, 4)
    ref = ((torch.softmax(torch.tanh(X @ W1) @ W2, -1) - T) ** 2).mean()
    ref.backward()
    (l,) = ex.step({"x": X, "t": T})
    assert abs(float(l) - float(ref)) < 1e-5
    assert torch.allclose(ex.store.master_view(w1.node), W1.detach() - 0.5 * W1.grad, atol=1e-5)
    assert torch.allclose(ex.store.master_view(w2.node), W2.detach() - 0.5 * W2.grad, atol=1e-5)


def test_fx_trace_matches_eager_torch_training():
    """torch.fx client frontend: traced nn.Module trains identically to eager PyTorch + torch.optim.SGD."""
    import torch.nn as nn
    from tepdist_b200.frontend.trace import trace

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(16, 32)
            self.ln = nn.LayerNorm(32)
            self.fc2 = nn.Linear(32, 8)

        def forward(self, x):
            return self.fc2(torch.tanh(self.ln(self.fc1(x))))

    torch.manual_seed(0)
    net = Net()
    x, y = torch.randn(8, 16), torch.randn(8, 8)
    tr = trace(net, {"x": x}, loss="mse", label_name="t", label_example=y, optimizer="sgd", lr=0.1)
    ex = Executor(tr.graph, torch.device("cpu"))
    tr.load_state_dict_into(ex, net.state_dict())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for _ in range(3):
        ref = ((net(x) - y) ** 2).mean()
        opt.zero_grad(); ref.backward(); opt.step()
        (l,) = ex.step({"x": x, "t": y})
        assert abs(float(l) - float(ref)) < 1e-5
    assert torch.allclose(ex.store.state_dict()["fc1/weight"], net.fc1.weight.detach(), atol=1e-5)


@pytest.mark.parametrize("make", [
    lambda ps: torch.optim.AdamW(ps, lr=0.01, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1),
    lambda ps: torch.optim.Adam(ps, lr=0.01, betas=(0.8, 0.99), eps=1e-7),
    lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9),
    lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, nesterov=True),
], ids=["adamw", "adam", "momentum", "nesterov"])
def test_fx_trace_takes_the_users_torch_optimizer(make):
    """`trace(..., optimizer=<torch.optim instance>)`: the traced step trains like eager PyTorch stepping that very optimizer."""
    import torch.nn as nn
    from tepdist_b200.frontend.trace import trace

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.ln, self.fc2 = nn.Linear(16, 32), nn.LayerNorm(32), nn.Linear(32, 8)

        def forward(self, x):
            return self.fc2(torch.tanh(self.ln(self.fc1(x))))

    torch.manual_seed(0)
    net = Net()
    x, y = torch.randn(8, 16), torch.randn(8, 8)
    opt = make(net.parameters())
    tr = trace(net, {"x": x}, loss="mse", label_name="t", label_example=y, optimizer=opt)
    ex = Executor(tr.graph, torch.device("cpu"))
    tr.load_state_dict_into(ex, net.state_dict())
    for _ in range(4):
        ref = ((net(x) - y) ** 2).mean()
        opt.zero_grad(); ref.backward(); opt.step()
        (l,) = ex.step({"x": x, "t": y})
        assert abs(float(l) - float(ref)) < 2e-5
    for name, p in net.named_parameters():
        assert torch.allclose(ex.store.state_dict()[name.replace(".", "/")], p.detach(), atol=2e-5), name


def test_fx_trace_rejects_optimizer_settings_it_cannot_reproduce():
    import torch.nn as nn
    from tepdist_b200.frontend.trace import optimizer_from_torch
    ps = list(nn.Linear(4, 4).parameters())
    for bad in (torch.optim.Adam(ps, weight_decay=0.1), torch.optim.SGD(ps, lr=0.1, weight_decay=0.1), torch.optim.RMSprop(ps),
                torch.optim.AdamW([{"params": ps[:1], "weight_decay": 0.0}, {"params": ps[1:], "weight_decay": 0.1}])):
        with pytest.raises(NotImplementedError):
            optimizer_from_torch(bad)


def test_fx_trace_transformer_block_with_sdpa():
    """A torch transformer block written with F.scaled_dot_product_attention traces onto the fused attention op and
    trains like eager PyTorch (chunk / getitem / view / transpose plumbing included)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from tepdist_b200.frontend.trace import trace

    class Block(nn.Module):
        def __init__(self, C=32, H=2, V=50):
            super().__init__()
            self.H = H
            self.emb = nn.Embedding(V, C)
            self.ln1, self.ln2 = nn.LayerNorm(C), nn.LayerNorm(C)
            self.qkv, self.proj = nn.Linear(C, 3 * C), nn.Linear(C, C)
            self.fc, self.out = nn.Linear(C, 4 * C), nn.Linear(4 * C, C)
            self.head = nn.Linear(C, V, bias=False)

        def forward(self, tokens):
            x = self.emb(tokens)
            B, S, C = 2, 16, 32
            q, k, v = self.qkv(self.ln1(x)).chunk(3, dim=-1)
            q, k, v = (t.view(B, S, self.H, C // self.H).transpose(1, 2) for t in (q, k, v))
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, C)
            x = x + self.proj(a)
            x = x + self.out(F.gelu(self.fc(self.ln2(x)), approximate="tanh"))
            return self.head(x)

    torch.manual_seed(0)
    net = Block()
    tok = torch.randint(0, 50, (2, 16))
    lab = torch.randint(0, 50, (2, 16))
    tr = trace(net, {"tokens": tok}, loss="cross_entropy", label_example=lab.int(), optimizer="sgd", lr=0.1)
    assert any(n.op == "attention" for n in tr.graph.nodes)
    ex = Executor(tr.graph, torch.device("cpu"))
    tr.load_state_dict_into(ex, net.state_dict())
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for _ in range(3):
        ref = F.cross_entropy(net(tok).view(-1, 50), lab.view(-1))
        opt.zero_grad(); ref.backward(); opt.step()
        (l,) = ex.step({"tokens": tok.int(), "labels": lab.int()})
        assert abs(float(l) - float(ref)) < 2e-4, (float(l), float(ref))
