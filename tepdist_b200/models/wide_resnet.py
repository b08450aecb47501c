"""Wide-ResNet (reference examples/wide_resnet: bottleneck ResNet-50/101 with widened channels, 7 sizes 250M..13B,
fake ImageNet input, Adam) in planner IR.  Convolutions / batch-norm execute through cuDNN exactly as the reference's
kConvolution custom-calls do; the planner sees batch / in-channel / out-channel split proposals."""
from __future__ import annotations

from dataclasses import dataclass

from ..frontend.builder import GraphBuilder, build_training_step
from ..ir import Graph, Value

# (base channels C, width factor W, blocks per stack) — examples/wide_resnet/train_imagenet.py:13-22
WRESNET_SPECS = [
    (160, 2, [3, 4, 6, 3]),    # 250M (50 layers)
    (224, 2, [3, 4, 6, 3]),    # 500M
    (320, 2, [3, 4, 6, 3]),    # 1B
    (448, 2, [3, 4, 6, 3]),    # 2B
    (640, 2, [3, 4, 6, 3]),    # 4B
    (320, 16, [3, 4, 6, 3]),   # 6.8B
    (320, 12, [3, 4, 23, 3]),  # 13B (101 layers)
]


@dataclass
class WideResNetConfig:
    model_type: int = 1
    batch: int = 4
    image: int = 224
    classes: int = 1000
    lr: float = 0.1

    @property
    def spec(self):
        return WRESNET_SPECS[self.model_type]


def _conv_bn(b: GraphBuilder, x: Value, cin: int, cout: int, k: int, stride: int, name: str, relu: bool = True) -> Value:
    import math
    w = b.parameter(f"{name}/w", (cout, cin, k, k), {"kind": "truncated_normal", "std": math.sqrt(2.0 / (cin * k * k))})
    g = b.parameter(f"{name}/bn/g", (cout,), {"kind": "constant", "value": 1.0})
    bt = b.parameter(f"{name}/bn/b", (cout,), {"kind": "constant", "value": 0.0})
    y = b.batchnorm(b.conv2d(x, w, stride, k // 2, name=name), g, bt, name=f"{name}/bn")
    return b.relu(y, name=f"{name}/relu") if relu else y


def _block(b: GraphBuilder, x: Value, cin: int, f: int, width: int, stride: int, name: str):
    y = _conv_bn(b, x, cin, f, 1, 1, f"{name}/a")
    y = _conv_bn(b, y, f, f * width, 3, stride, f"{name}/b")
    y = _conv_bn(b, y, f * width, 4 * f, 1, 1, f"{name}/c", relu=False)
    sc = x
    if cin != 4 * f or stride != 1:
        sc = _conv_bn(b, x, cin, 4 * f, 1, stride, f"{name}/shortcut", relu=False)
    return b.relu(b.add(y, sc, name=f"{name}/add"), name=f"{name}/out"), 4 * f


def build_wide_resnet_graph(cfg: WideResNetConfig, dtype: str = "bf16", optimizer: str = "adamw") -> Graph:
    C, W, blocks = cfg.spec
    b = GraphBuilder(f"wide_resnet_{cfg.model_type}", compute_dtype=dtype)
    x = b.input("images", (cfg.batch, 3, cfg.image, cfg.image), dtype)
    labels = b.input("labels", (cfg.batch,), "i32")
    y = _conv_bn(b, x, 3, 64, 7, 2, "conv1")
    y = b.maxpool2d(y, 3, 2, 1, name="pool1")
    cin = 64
    for i, nb in enumerate(blocks):
        f = C * (2 ** i)
        for j in range(nb):
            y, cin = _block(b, y, cin, f, W, (1 if i == 0 else 2) if j == 0 else 1, f"scale{i + 2}/block{j + 1}")
    feat = b.global_avgpool(y, name="avg_pool")
    wf = b.parameter("fc/w", (cfg.classes, cin), {"kind": "truncated_normal", "std": 0.01})
    bf = b.parameter("fc/b", (cfg.classes,), {"kind": "constant", "value": 0.0})
    logits = b.linear(feat, wf, bf, name="fc")
    loss = b.softmax_xent(logits, labels, vocab=cfg.classes, name="loss")
    # reference: train_imagenet.py:42 AdamOptimizer(0.1) (the default here), resnet_train.py:34 MomentumOptimizer(0.01, 0.9)
    g = build_training_step(b, loss, optimizer, lr=cfg.lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, momentum=0.9)
    g.meta["model"] = {"family": "wide_resnet", "model_type": cfg.model_type, "batch": cfg.batch, "image": cfg.image}
    return g
