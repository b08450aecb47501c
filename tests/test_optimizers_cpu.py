"""Optimizers of the reference's example suites (GPT-2 Adam / Adafactor; GPT-MoE AdamW / Adafactor / LAMB / SM3):
end-to-end numerics of graph autodiff + `apply_<kind>` nodes against torch autograd with torch.optim (SGD-momentum, Adam, AdamW)
or against plain re-statements of the update rules written here on whole tensors (LAMB, Adafactor, SM3 -- deliberately not
sharing code with tepdist_b200/runtime/optimizers.py).  The sharded variants (ZeRO chunks, tensor-parallel shards) are compared
with these single-process runs in test_distributed_cpu.py::test_reduction_optimizers_match_single_process_when_sharded."""
import math

import pytest
import torch

from tepdist_b200.frontend.builder import OPTIMIZERS, GraphBuilder, build_training_step
from tepdist_b200.runtime.executor import Executor


def build_mlp(opt, **hp):
    b = GraphBuilder("opt_mlp", compute_dtype="f32")
    x = b.input("x", (8, 16), "f32")
    t = b.input("t", (8, 4), "f32")
    w1 = b.parameter("w1", (16, 32), {"kind": "normal", "std": 0.3})
    b1 = b.parameter("b1", (32,), {"kind": "constant", "value": 0.1})
    b.g.nodes[b1.node].attrs["decay"] = False
    w2 = b.parameter("w2", (32, 4), {"kind": "normal", "std": 0.3})
    d = b.sub(b.matmul(b.tanh(b.add(b.matmul(x, w1, name="fc1"), b1)), w2, name="fc2"), t)
    return build_training_step(b, b.reduce_mean(b.mul(d, d), [0, 1], name="loss"), opt, **hp)


def torch_loss(p, x, t):
    return ((torch.tanh(x @ p["w1"] + p["b1"]) @ p["w2"] - t) ** 2).mean()


# ------------------------------------------------------------------------------------------ reference update rules
class RefLamb:
    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, no_decay=()):
        self.p, self.hp, self.t, self.no_decay = params, (lr, beta1, beta2, eps, weight_decay), 0, set(no_decay)
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads):
        lr, b1, b2, eps, wd = self.hp
        self.t += 1
        for k, p in self.p.items():
            g = grads[k]
            self.m[k] = b1 * self.m[k] + (1 - b1) * g
            self.v[k] = b2 * self.v[k] + (1 - b2) * g * g
            u = (self.m[k] / (1 - b1 ** self.t)) / ((self.v[k] / (1 - b2 ** self.t)).sqrt() + eps)
            ratio = 1.0
            if k not in self.no_decay:
                u = u + wd * p
                wn, un = float(p.norm()), float(u.norm())
                ratio = wn / un if wn > 0 and un > 0 else 1.0
            p.sub_(lr * ratio * u)


class RefAdafactor:
    def __init__(self, params, lr=None, clip=1.0, eps1=1e-30, eps2=1e-3):
        self.p, self.lr, self.clip, self.eps1, self.eps2, self.t = params, lr, clip, eps1, eps2, 0
        self.vr = {k: torch.zeros(v.shape[:-1]) for k, v in params.items() if v.dim() >= 2}
        self.vc = {k: torch.zeros(v.shape[:-2] + v.shape[-1:]) for k, v in params.items() if v.dim() >= 2}
        self.v = {k: torch.zeros_like(v) for k, v in params.items() if v.dim() < 2}

    def step(self, grads):
        self.t += 1
        beta2 = 1.0 - self.t ** -0.8
        for k, p in self.p.items():
            g2 = grads[k] ** 2 + self.eps1
            if p.dim() >= 2:
                self.vr[k] = beta2 * self.vr[k] + (1 - beta2) * g2.mean(-1)
                self.vc[k] = beta2 * self.vc[k] + (1 - beta2) * g2.mean(-2)
                vhat = (self.vr[k] / self.vr[k].mean(-1, keepdim=True)).unsqueeze(-1) * self.vc[k].unsqueeze(-2)
            else:
                self.v[k] = beta2 * self.v[k] + (1 - beta2) * g2
                vhat = self.v[k]
            u = grads[k] / vhat.sqrt()
            u = u / max(1.0, float((u ** 2).mean().sqrt()) / self.clip)
            lr = self.lr if self.lr is not None else min(1e-2, 1.0 / math.sqrt(self.t))
            p.sub_(lr * max(self.eps2, float((p ** 2).mean().sqrt())) * u)


class RefSM3:
    def __init__(self, params, lr, momentum=0.0):
        self.p, self.lr, self.mu = params, lr, momentum
        self.acc = {k: ([torch.zeros(d) for d in v.shape] if v.dim() > 1 else [torch.zeros_like(v)]) for k, v in params.items()}
        self.mom = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads):
        for k, p in self.p.items():
            g = grads[k]
            if p.dim() > 1:
                nu = torch.full_like(p, float("inf"))
                for i, a in enumerate(self.acc[k]):
                    shape = [1] * p.dim()
                    shape[i] = -1
                    nu = torch.minimum(nu, a.reshape(shape).expand_as(p))
                nu = nu + g * g
                for i in range(p.dim()):
                    self.acc[k][i] = nu.amax(dim=[j for j in range(p.dim()) if j != i])
            else:
                self.acc[k][0] = self.acc[k][0] + g * g
                nu = self.acc[k][0]
            u = (1 - self.mu) * g / (nu + 1e-30).sqrt()
            if self.mu > 0:
                self.mom[k] = self.mu * self.mom[k] + u
                u = self.mom[k]
            p.sub_(self.lr * u)


class TorchOptim:
    def __init__(self, params, make):
        self.p = params
        self.leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.opt = make(self.leaves)

    def step(self, grads):
        for k, v in self.leaves.items():
            v.grad = grads[k].clone()
        self.opt.step()
        for k in self.p:
            self.p[k].copy_(self.leaves[k].detach())


CASES = {
    "sgd": (dict(lr=0.1), lambda p: TorchOptim(p, lambda l: torch.optim.SGD(l.values(), lr=0.1))),
    "momentum": (dict(lr=0.05, momentum=0.9), lambda p: TorchOptim(p, lambda l: torch.optim.SGD(l.values(), lr=0.05, momentum=0.9))),
    "adam": (dict(lr=0.01, beta1=0.9, beta2=0.98, eps=1e-9),
             lambda p: TorchOptim(p, lambda l: torch.optim.Adam(l.values(), lr=0.01, betas=(0.9, 0.98), eps=1e-9))),
    "adamw": (dict(lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.1),
              lambda p: TorchOptim(p, lambda l: torch.optim.AdamW(
                  [{"params": [l["w1"], l["w2"]], "weight_decay": 0.1}, {"params": [l["b1"]], "weight_decay": 0.0}],
                  lr=0.01, betas=(0.9, 0.999), eps=1e-8))),
    "lamb": (dict(lr=0.01, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.05),
             lambda p: RefLamb(p, 0.01, 0.9, 0.999, 1e-6, 0.05, no_decay=("b1",))),
    "adafactor": (dict(lr=0.05), lambda p: RefAdafactor(p, lr=0.05)),
    "adafactor_relative_step": (dict(lr=None), lambda p: RefAdafactor(p, lr=None)),
    "sm3": (dict(lr=0.1), lambda p: RefSM3(p, 0.1)),
    "sm3_momentum": (dict(lr=0.1, momentum=0.9), lambda p: RefSM3(p, 0.1, 0.9)),
}


def test_every_builder_optimizer_has_a_numerics_case():
    assert {c.split("_")[0] for c in CASES} == set(OPTIMIZERS)


@pytest.mark.parametrize("case", sorted(CASES))
def test_optimizer_matches_reference_update_rule(case):
    hp, make_ref = CASES[case]
    kind = case.split("_")[0]
    ex = Executor(build_mlp(kind, **hp), torch.device("cpu"), seed=5, use_cuda_graph=False)
    params = {k: v.clone() for k, v in ex.store.state_dict().items() if k in ("w1", "b1", "w2")}
    ref = make_ref(params)
    torch.manual_seed(1)
    losses = []
    for step in range(6):
        feeds = {"x": torch.randn(8, 16), "t": torch.randn(8, 4)}
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        loss = torch_loss(leaves, feeds["x"], feeds["t"])
        loss.backward()
        got = float(ex.step(feeds)[0])
        losses.append(got)
        assert got == pytest.approx(float(loss.detach()), rel=2e-5, abs=1e-6), (case, step)
        with torch.no_grad():
            ref.step({k: v.grad for k, v in leaves.items()})
        mine = ex.store.state_dict()
        for k in params:
            assert torch.allclose(mine[k], params[k], rtol=2e-4, atol=2e-6), (case, step, k, float((mine[k] - params[k]).abs().max()))
    assert all(math.isfinite(l) for l in losses)


def test_reduced_shape_slots_are_saved_and_restored():
    """Adafactor / SM3 slots live outside the flat m / v buffers; state_dict round-trips them and training continues identically."""
    for kind, hp in (("adafactor", dict(lr=0.05)), ("sm3", dict(lr=0.1, momentum=0.9))):
        torch.manual_seed(2)
        feeds = [{"x": torch.randn(8, 16), "t": torch.randn(8, 4)} for _ in range(6)]
        a = Executor(build_mlp(kind, **hp), torch.device("cpu"), seed=5, use_cuda_graph=False)
        for f in feeds[:3]:
            a.step(f)
        sd = a.store.state_dict()
        names = set(sd)
        assert ({"w1/vr", "w1/vc", "b1/vf"} <= names) if kind == "adafactor" else ({"w1/acc0", "w1/acc1", "b1/acc", "w1/mom"} <= names)
        b = Executor(build_mlp(kind, **hp), torch.device("cpu"), seed=99, use_cuda_graph=False)
        b.store.load_state_dict(sd)
        b.step_count = a.step_count
        for f in feeds[3:]:
            assert float(a.step(f)[0]) == float(b.step(f)[0])


@pytest.mark.parametrize("kind", ["sgd", "momentum", "adamw", "lamb"])
def test_learning_rate_schedule_matches_torch_lambda_lr(kind):
    """Trainer.set_lr_schedule: warm-up + cosine decay evaluated per step; the same trajectory as torch.optim + LambdaLR
    (LAMB: against the reference update rule with the rate set by hand)."""
    from tepdist_b200.api import Trainer
    from tepdist_b200.utils.schedules import warmup_cosine, warmup_linear_decay, rsqrt_decay
    hp, make_ref = CASES[kind]
    base = hp["lr"]
    sched = warmup_cosine(base, warmup_steps=3, total_steps=8, end_ratio=0.1)
    assert sched(1) == pytest.approx(base / 3) and sched(3) == pytest.approx(base) and sched(8) == pytest.approx(0.1 * base)
    assert sched(100) == pytest.approx(0.1 * base)
    lin = warmup_linear_decay(1.0, 2, 6)
    assert [round(lin(s), 3) for s in (1, 2, 4, 6, 9)] == [0.5, 1.0, 0.5, 0.0, 0.0]
    assert rsqrt_decay(1.0, 4)(2) == 0.5 and rsqrt_decay(1.0, 4)(16) == 0.5
    tr = Trainer(build_mlp(kind, **hp), device=torch.device("cpu"), use_cuda_graph=False, seed=5)
    tr.set_lr_schedule(sched)
    params = {k: v.clone() for k, v in tr.exec.store.state_dict().items() if k in ("w1", "b1", "w2")}
    ref = make_ref(params)
    torch.manual_seed(3)
    for step in range(1, 9):
        feeds = {"x": torch.randn(8, 16), "t": torch.randn(8, 4)}
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        torch_loss(leaves, feeds["x"], feeds["t"]).backward()
        if isinstance(ref, TorchOptim):
            for gr in ref.opt.param_groups:
                gr["lr"] = sched(step)
        else:
            ref.hp = (sched(step),) + tuple(ref.hp[1:])
        tr.step(feeds)
        with torch.no_grad():
            ref.step({k: v.grad for k, v in leaves.items()})
        mine = tr.exec.store.state_dict()
        for k in params:
            assert torch.allclose(mine[k], params[k], rtol=2e-4, atol=2e-6), (kind, step, k)
    tr.set_lr_schedule(None)
    assert tr.exec._lr(123.0) == hp["lr"]


def test_schedule_spec_travels_with_the_serialized_graph():
    """`build_training_step(..., schedule={...})`: the spec sits in graph.meta["optimizer"], survives JSON serialization (what a
    client sends to a server) and drives the executor's learning rate."""
    import json
    from tepdist_b200.ir import Graph
    from tepdist_b200.utils.schedules import from_spec, warmup_cosine
    spec = {"kind": "warmup_cosine", "warmup_steps": 2, "total_steps": 6, "end_ratio": 0.1}
    g = build_mlp("adamw", lr=0.01, weight_decay=0.0, schedule=spec)
    g2 = Graph.from_dict(json.loads(json.dumps(g.to_dict())))
    assert g2.meta["optimizer"]["schedule"] == spec
    want = warmup_cosine(0.01, 2, 6, 0.1)
    assert [from_spec(spec, 0.01)(s) for s in range(1, 8)] == [want(s) for s in range(1, 8)]
    ex = Executor(g2, torch.device("cpu"), seed=5, use_cuda_graph=False)
    feeds = {"x": torch.randn(8, 16), "t": torch.randn(8, 4)}
    seen = []
    for _ in range(7):
        ex.step(feeds)
        seen.append(float(ex.hyper[0]))
    assert seen == pytest.approx([want(s) for s in range(1, 8)], rel=1e-6)
    with pytest.raises(ValueError):
        from_spec({"kind": "nope"}, 0.1)


@pytest.mark.parametrize("mode", ["global", "local"])
@pytest.mark.parametrize("kind", ["adamw", "sgd", "lamb"])
def test_gradient_clipping_matches_torch(kind, mode):
    """clip_norm='global' == torch.nn.utils.clip_grad_norm_ over all gradients, 'local' == the same per tensor (reference
    examples/gpt_moe/optimizers/__init__.py:150-159), applied before the update of every optimizer."""
    hp, make_ref = CASES[kind]
    c = 0.05
    ex = Executor(build_mlp(kind, clip_norm=mode, clip_norm_value=c, **hp), torch.device("cpu"), seed=5, use_cuda_graph=False)
    assert ex.clip == (mode, c) and not ex.fused_apply_ok
    params = {k: v.clone() for k, v in ex.store.state_dict().items() if k in ("w1", "b1", "w2")}
    ref = make_ref(params)
    torch.manual_seed(1)
    clipped_once = False
    for step in range(5):
        feeds = {"x": torch.randn(8, 16), "t": torch.randn(8, 4)}
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        torch_loss(leaves, feeds["x"], feeds["t"]).backward()
        grads = {k: v.grad for k, v in leaves.items()}
        if mode == "global":
            total = torch.sqrt(sum((g_ * g_).sum() for g_ in grads.values()))
            clipped_once |= bool(total > c)
            grads = {k: g_ * min(1.0, c / (float(total) + 1e-6)) for k, g_ in grads.items()}
        else:
            clipped_once |= any(bool(g_.norm() > c) for g_ in grads.values())
            grads = {k: g_ * min(1.0, c / (float(g_.norm()) + 1e-6)) for k, g_ in grads.items()}
        ex.step(feeds)
        with torch.no_grad():
            ref.step(grads)
        mine = ex.store.state_dict()
        for k in params:
            assert torch.allclose(mine[k], params[k], rtol=2e-4, atol=2e-6), (kind, mode, step, k)
    assert clipped_once, "the threshold never bit: the test would prove nothing"
    with pytest.raises(ValueError):
        Executor(build_mlp(kind, clip_norm="sideways", **hp), torch.device("cpu"), use_cuda_graph=False)
