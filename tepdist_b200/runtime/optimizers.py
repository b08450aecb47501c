"""Per-variable optimizer updates beyond SGD / AdamW: momentum, LAMB, Adafactor, SM3 -- the optimizers of the reference's example
suites (examples/GPT2/optimizers.py:102 AdafactorOptimizer; examples/gpt_moe/optimizers/{adafactor,lamb_weight_decay_optimizer,
sm3}.py), which there are TF graphs of primitive ops that the planner shards op by op.

Here each is ONE `apply_<kind>` node acting on this rank's VIEW of the variable (the whole variable, a ZeRO chunk, a tensor-
parallel shard or several of these nested).  What makes them different from AdamW is that they reduce over the variable
(norms, row / column means, per-dimension maxima); `Shards` describes how the view was cut out of the full variable and completes
those reductions across the ranks that hold the other pieces, so a sharded update is bit-for-bit the update of the whole
variable up to summation order.

All step-dependent scalars come from the device tensor `hyper` (see Executor._set_hyper), never from Python floats, so the
update can be captured in a CUDA graph.  All math is fp32 on the master copy.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

# slots of `hyper` (fp32 device tensor, filled on the host once per step)
H_LR, H_BC1, H_BC2, H_GRAD_SCALE, H_AF_DECAY, H_AF_REL_LR = range(6)
HYPER_SIZE = 8


class Shards:
    """How this rank's view was cut out of the full variable: [(dim, num, level)], outermost first.  `reduce(t, dims, op)`
    completes a reduction that ran over the local extent of `dims` (None = all dims) across every level that shards one of
    them."""

    def __init__(self, chain: Sequence[Tuple[int, int, int]] = (),
                 all_reduce: Optional[Callable[[torch.Tensor, int, str], None]] = None):
        self.chain = list(chain)
        self._ar = all_reduce

    def factor(self, dims: Optional[Sequence[int]] = None) -> int:
        f = 1
        for d, num, _ in self.chain:
            if dims is None or d in dims:
                f *= num
        return f

    def reduce(self, t: torch.Tensor, dims: Optional[Sequence[int]], op: str = "sum") -> torch.Tensor:
        if self._ar is not None:
            for d, num, lvl in self.chain:
                if num > 1 and (dims is None or d in dims):
                    self._ar(t, lvl, op)
        return t


def _rms(x: torch.Tensor, sh: Shards) -> torch.Tensor:
    s = sh.reduce((x * x).sum().reshape(1), None)
    return torch.sqrt(s / float(x.numel() * sh.factor(None))).reshape(())


def momentum_step(p, g, slots, hp, hyper, sh: Shards, decay: bool) -> None:
    (v,) = slots
    v.mul_(hp.get("momentum", 0.9)).add_(g)
    upd = g + hp.get("momentum", 0.9) * v if hp.get("nesterov", False) else v
    p.sub_(hyper[H_LR] * upd)


def lamb_step(p, g, slots, hp, hyper, sh: Shards, decay: bool) -> None:
    m, v = slots
    b1, b2, eps = hp.get("beta1", 0.9), hp.get("beta2", 0.999), hp.get("eps", 1e-6)
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    upd = (m / hyper[H_BC1]) / (torch.sqrt(v / hyper[H_BC2]) + eps)
    wd = hp.get("weight_decay", 0.0)
    if decay and wd:
        upd = upd + wd * p
    ratio = 1.0
    if decay or hp.get("adapt_all", False):       # layer adaptation is skipped for the variables excluded from weight decay
        sq = sh.reduce(torch.stack([(p * p).sum(), (upd * upd).sum()]), None)
        wn, un = torch.sqrt(sq[0]), torch.sqrt(sq[1])
        one = torch.ones_like(wn)
        ratio = torch.where(wn > 0, torch.where(un > 0, wn / un, one), one)
    p.sub_(hyper[H_LR] * ratio * upd)


def adafactor_step(p, g, slots, hp, hyper, sh: Shards, decay: bool) -> None:
    eps1, eps2, clip = hp.get("eps1", 1e-30), hp.get("eps2", 1e-3), hp.get("clipping_threshold", 1.0)
    beta2 = hyper[H_AF_DECAY] if hp.get("decay_rate") is None else hp["decay_rate"]
    g2 = g * g + eps1
    r = p.dim()
    if len(slots) == 2:
        vr, vc = slots
        row = sh.reduce(g2.sum(-1), [r - 1]) / float(p.shape[-1] * sh.factor([r - 1]))       # mean over columns
        col = sh.reduce(g2.sum(-2), [r - 2]) / float(p.shape[-2] * sh.factor([r - 2]))       # mean over rows
        vr.mul_(beta2).add_(row * (1.0 - beta2))
        vc.mul_(beta2).add_(col * (1.0 - beta2))
        row_mean = sh.reduce(vr.sum(-1, keepdim=True), [r - 2]) / float(p.shape[-2] * sh.factor([r - 2]))
        upd = g * torch.rsqrt(vr / row_mean).unsqueeze(-1) * torch.rsqrt(vc).unsqueeze(-2)
    else:
        (vf,) = slots
        vf.mul_(beta2).add_(g2 * (1.0 - beta2))
        upd = g * torch.rsqrt(vf)
    if clip is not None:
        upd = upd / torch.clamp(_rms(upd, sh) / clip, min=1.0)
    lr = hyper[H_AF_REL_LR] if hp.get("lr") is None else hyper[H_LR]
    if hp.get("multiply_by_parameter_scale", True):
        lr = lr * torch.clamp(_rms(p, sh), min=eps2)
    wd = hp.get("weight_decay", 0.0)
    if decay and wd:
        upd = upd + wd * p
    p.sub_(lr * upd)


def sm3_step(p, g, slots, hp, hyper, sh: Shards, decay: bool) -> None:
    mu = hp.get("momentum", 0.0)
    mom = slots[-1] if mu > 0 else None
    accs = slots[:-1] if mu > 0 else slots
    r = p.dim()
    if r > 1:
        nu = None
        for i, a in enumerate(accs):
            b = a.reshape([-1 if j == i else 1 for j in range(r)])
            nu = b if nu is None else torch.minimum(nu, b)
        nu = nu + g * g
        for i, a in enumerate(accs):
            others = [j for j in range(r) if j != i]
            a.copy_(sh.reduce(nu.amax(dim=others), others, "max"))
    else:
        (a,) = accs
        a.add_(g * g)
        nu = a
    upd = (1.0 - mu) * g * torch.rsqrt(nu + 1e-30)
    if mom is not None:
        mom.mul_(mu).add_(upd)
        upd = mom
    p.sub_(hyper[H_LR] * upd)


STEP: Dict[str, Callable] = {"apply_momentum": momentum_step, "apply_lamb": lamb_step, "apply_adafactor": adafactor_step,
                             "apply_sm3": sm3_step}


def host_hyper(opt: Dict, step: int, lr: float) -> List[float]:
    """Values of `hyper` for optimizer step `step` (1-based)."""
    b1, b2 = opt.get("beta1", 0.9), opt.get("beta2", 0.999)
    t = max(1, step)
    return [lr, 1.0 - b1 ** t, 1.0 - b2 ** t, 1.0,
            1.0 - float(t) ** (-opt.get("decay_pow", 0.8)),          # Adafactor: adafactor_decay_rate_pow(0.8)
            min(1e-2, 1.0 / float(t) ** 0.5), 0.0, 0.0]              # Adafactor: relative step size when no lr is given
