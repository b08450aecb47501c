"""Python face of the C++ planner (tepdist_b200._C)."""
from __future__ import annotations

from typing import Any, Dict

from ..ir import Graph


def to_native(g: Graph):
    """ir.Graph -> _C.Graph (same node ids)."""
    from .. import _C
    cg = _C.Graph()
    cg.name = g.name
    for n in g.nodes:
        attrs: Dict[str, Any] = {}
        for k, v in n.attrs.items():
            if k == "sharding":  # {"0": {"dim": d, "num": n}} -> flat annotation on output 0
                s = v.get("0")
                if s is not None:
                    attrs["shard_dim"] = int(s["dim"])
                    attrs["shard_num"] = int(s["num"])
                continue
            attrs[k] = v
        nid = cg.add_node(n.op, [(v.node, v.idx) for v in n.inputs], [(list(t.shape), t.dtype) for t in n.outputs], attrs,
                          n.name, n.group, n.backward)
        if n.stage >= 0:      # (a planned graph converted back keeps its pipeline stages; from_native reads them)
            cg.set_node_stage(n.id if nid is None else nid, n.stage)
    cg.set_outputs([(v.node, v.idx) for v in g.outputs])
    for var, v in g.updates.items():
        cg.set_update(var, v.node, v.idx)
    return cg


def from_native(cg) -> Graph:
    """_C.Graph -> ir.Graph (node ids preserved)."""
    from ..ir import Node, TensorType, Value
    g = Graph(cg.name)
    for i in range(cg.num_nodes()):
        outs = [TensorType(tuple(s), d) for s, d in cg.node_outputs(i)]
        n = Node(i, cg.node_op(i), [Value(a, b) for a, b in cg.node_inputs(i)], outs, dict(cg.node_attrs(i)),
                 cg.node_name(i), cg.node_group(i), cg.node_backward(i), cg.node_stage(i))
        g.nodes.append(n)
        g._next_group = max(g._next_group, n.group + 1)
    g.outputs = [Value(a, b) for a, b in cg.outputs()]
    g.updates = {k: Value(a, b) for k, a, b in cg.updates()}
    g.meta = dict(cg.meta)
    return g


def merge_client_attrs(dst: Graph, src: Graph) -> None:
    """Nested attrs (initialiser specs) never cross into C++; copy them back by node name."""
    by_name = {n.name: n for n in src.nodes}
    for n in dst.nodes:
        s = by_name.get(n.name)
        if s is not None and n.op == s.op:
            for k, v in s.attrs.items():
                if isinstance(v, dict) and k not in n.attrs:
                    n.attrs[k] = v
    dst.meta.update({k: v for k, v in src.meta.items() if k not in dst.meta})


def liveness_optimize(g: Graph, min_bytes: int = 1 << 20):
    """B6 HloLivenessOptimizer (reference hlo_liveness_optimizer.cc:26-54; C++: transform.cc LivenessOptimize): give every user of
    a large convert(parameter) its own copy of the convert, placed right in front of it, so a casted weight is not kept alive
    from the forward to the backward pass.  Graphs built by models/ keep variables in compute precision in the flat store and
    contain no such converts; graphs from the builder / torch.fx frontend with explicit mixed-precision casts do.
    Returns (graph, number of copies); the input graph is returned unchanged when there is nothing to do."""
    from .. import _C
    users: Dict[int, set] = {}
    for n in g.nodes:
        for v in n.inputs:
            src = g.nodes[v.node]
            if src.op == "cast" and g.nodes[src.inputs[0].node].op == "parameter":
                users.setdefault(src.id, set()).add(n.id)
    if not any(len(u) > 1 for u in users.values()):
        return g, 0
    cg = to_native(g)
    copies = int(_C.liveness_optimize(cg, int(min_bytes)))
    if not copies:
        return g, 0
    out = from_native(cg)
    merge_client_attrs(out, g)
    return out, copies
