"""Planner IR — the typed, static-shape graph that replaces HLO in this system.

The PyTorch client lowers a whole training step (forward, backward, optimizer update) into this IR
(``frontend/``), ships it to the C++ planner (``csrc/``, same schema) which annotates every value with
a DistSpec, rewrites it into per-shard graphs with collectives, cuts it into micro-batch / pipeline
sub-graphs, and the runtime executes the result.

Reference parity: HloModule / HloInstruction + the TePDist metadata extensions
(`op_group`, `backward`: tensorflow/compiler/xla/xla_data.proto diff fields 5-6; variable_map /
init_specs / fetch list: xla/service/hlo_module.h diff — SURVEY §2.C C3, C4).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

DTYPE_BYTES = {"bf16": 2, "f16": 2, "f32": 4, "i32": 4, "i64": 8, "bool": 1, "f8e4m3": 1}


@dataclass(frozen=True)
class Value:
    """One output of a node."""
    node: int
    idx: int = 0

    def key(self) -> Tuple[int, int]:
        return (self.node, self.idx)


@dataclass
class TensorType:
    shape: Tuple[int, ...]
    dtype: str

    def numel(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def nbytes(self) -> int:
        return self.numel() * DTYPE_BYTES[self.dtype]


@dataclass
class Node:
    id: int
    op: str
    inputs: List[Value]
    outputs: List[TensorType]
    attrs: Dict[str, Any] = field(default_factory=dict)
    name: str = ""
    group: int = -1        # op_group: a forward op and the gradient ops derived from it share an id
    backward: bool = False
    stage: int = -1        # pipeline stage (filled by the planner)

    def out(self, idx: int = 0) -> Value:
        return Value(self.id, idx)


# ops that introduce data into the step
SOURCE_OPS = ("parameter", "input", "constant", "state")
# ops that only move / reduce data across devices (inserted by the transforms, never by the client)
COLLECTIVE_OPS = ("all_reduce", "all_gather", "reduce_scatter", "all_to_all", "dynamic_slice", "send", "recv")


class Graph:
    """A whole training step.  Nodes are kept in a valid topological order (construction order)."""

    def __init__(self, name: str = "step"):
        self.name = name
        self.nodes: List[Node] = []
        self.outputs: List[Value] = []          # fetches (loss, metrics)
        self.updates: Dict[int, Value] = {}     # parameter/state node id -> updated value (in/out alias)
        self.meta: Dict[str, Any] = {}
        self._next_group = 0

    # ------------------------------------------------------------------ construction
    def add(self, op: str, inputs: Sequence[Value], outputs: Sequence[TensorType], attrs: Optional[Dict[str, Any]] = None,
            name: str = "", group: int = -1, backward: bool = False) -> Node:
        n = Node(len(self.nodes), op, list(inputs), list(outputs), dict(attrs or {}), name or f"{op}_{len(self.nodes)}",
                 group, backward)
        self.nodes.append(n)
        return n

    def new_group(self) -> int:
        g = self._next_group
        self._next_group += 1
        return g

    def type_of(self, v: Value) -> TensorType:
        return self.nodes[v.node].outputs[v.idx]

    def shape_of(self, v: Value) -> Tuple[int, ...]:
        return self.type_of(v).shape

    # ------------------------------------------------------------------ queries
    def params(self) -> List[Node]:
        return [n for n in self.nodes if n.op == "parameter"]

    def inputs(self) -> List[Node]:
        return [n for n in self.nodes if n.op == "input"]

    def users(self) -> Dict[Tuple[int, int], List[Tuple[int, int]]]:
        """value key -> [(user node id, operand index)]"""
        u: Dict[Tuple[int, int], List[Tuple[int, int]]] = {}
        for n in self.nodes:
            for i, v in enumerate(n.inputs):
                u.setdefault(v.key(), []).append((n.id, i))
        return u

    def validate(self) -> None:
        for n in self.nodes:
            for v in n.inputs:
                assert v.node < n.id, f"node {n.id} ({n.op}) uses later node {v.node}"
                assert v.idx < len(self.nodes[v.node].outputs)

    # ------------------------------------------------------------------ (de)serialisation
    def to_dict(self) -> Dict[str, Any]:
        return {
            "name": self.name,
            "meta": self.meta,
            "nodes": [
                {
                    "id": n.id, "op": n.op, "name": n.name,
                    "inputs": [[v.node, v.idx] for v in n.inputs],
                    "outputs": [[list(t.shape), t.dtype] for t in n.outputs],
                    "attrs": n.attrs, "group": n.group, "backward": n.backward, "stage": n.stage,
                }
                for n in self.nodes
            ],
            "outputs": [[v.node, v.idx] for v in self.outputs],
            "updates": [[k, v.node, v.idx] for k, v in sorted(self.updates.items())],
        }

    def to_json(self) -> str:
        return json.dumps(self.to_dict())

    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "Graph":
        g = Graph(d.get("name", "step"))
        g.meta = d.get("meta", {})
        for nd in d["nodes"]:
            n = Node(nd["id"], nd["op"], [Value(a, b) for a, b in nd["inputs"]],
                     [TensorType(tuple(s), t) for s, t in nd["outputs"]], nd.get("attrs", {}), nd.get("name", ""),
                     nd.get("group", -1), nd.get("backward", False), nd.get("stage", -1))
            assert n.id == len(g.nodes)
            g.nodes.append(n)
            g._next_group = max(g._next_group, n.group + 1)
        g.outputs = [Value(a, b) for a, b in d.get("outputs", [])]
        g.updates = {k: Value(a, b) for k, a, b in d.get("updates", [])}
        return g

    @staticmethod
    def from_json(s: str) -> "Graph":
        return Graph.from_dict(json.loads(s))

    def summary(self) -> str:
        from collections import Counter
        c = Counter(n.op for n in self.nodes)
        return f"Graph({self.name}: {len(self.nodes)} nodes; " + ", ".join(f"{k}={v}" for k, v in c.most_common(12)) + ")"


def numel(shape: Iterable[int]) -> int:
    n = 1
    for d in shape:
        n *= d
    return n
