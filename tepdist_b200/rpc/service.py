"""Client/server RPC layer (gRPC, generic byte methods — no protoc needed).

Reference parity: the client (CPU-only process holding the model script) ships the
whole-step graph once (`BuildExecutionPlan`), the master plans it and dispatches the plan to the workers, then every
step the client sends the sample inputs (`TransferToServerHost`) and calls `ExecutePlan`; variables never leave the
servers unless fetched (`FetchResourceVars`); `DoRemoteSave` / `DoRemoteRestore` drive the sharded checkpoints.
Master -> worker traffic (the reference's TransferModuleAndDefCtx / DispatchPlan / TransferHostRawData /
ExecuteRemotePlan gRPCs) rides on the torch.distributed control plane of the server job: one process per GPU.
"""
from __future__ import annotations

import io
import threading
from concurrent import futures
from typing import Any, Dict, List, Optional

import grpc
import torch
import torch.distributed as dist

from ..ir import Graph

SERVICE = "tepdist.TePDistService"
METHODS = ["BuildExecutionPlan", "ExecutePlan", "TransferToServerHost", "FetchResourceVars", "DoRemoteSave",
           "DoRemoteRestore", "GetServerInfo", "Shutdown"]
_MAX = [("grpc.max_send_message_length", -1), ("grpc.max_receive_message_length", -1)]


def pack(obj: Any) -> bytes:
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def unpack(b: bytes) -> Any:
    """Decode a request / reply.  `weights_only=True`: the restricted unpickler accepts containers, scalars, strings, tensors,
    dtypes and sizes only -- a request can carry data, never code (the reference's protobuf messages have the same property;
    a full pickle load on an unauthenticated port would hand every client arbitrary code execution on all server ranks)."""
    return torch.load(io.BytesIO(b), weights_only=True)


class ExecutionPlanCache:
    """handle <-> built plan (reference xla/service/execution_plan_cache.*, D15)."""

    def __init__(self):
        self._plans: Dict[int, Any] = {}
        self._next = 1
        self._lock = threading.Lock()

    def insert(self, plan) -> int:
        with self._lock:
            h = self._next
            self._next += 1
            self._plans[h] = plan
            return h

    def get(self, handle: int):
        return self._plans[handle]


class ServiceImpl:
    """Runs on EVERY server rank; only the master (rank 0) is reachable from the client."""

    def __init__(self, strategy: str = "auto", device: Optional[torch.device] = None, ckpt_root: str = ".",
                 comm_mode: str = "fused", use_cuda_graph: bool = False):
        from ..api import init_distributed
        self.ctx = init_distributed()
        self.rank, self.world = self.ctx["rank"], self.ctx["world"]
        self.device = device
        self.strategy, self.comm_mode, self.use_cuda_graph = strategy, comm_mode, use_cuda_graph
        self.cache = ExecutionPlanCache()
        self.host_inputs: Dict[str, torch.Tensor] = {}     # registered sample inputs (latest step)
        self.fake_input_cache: Optional[Dict[str, torch.Tensor]] = None
        self.exec_lock = threading.Lock()                  # reference execute_plan_mutex_
        self.ckpt_root = ckpt_root
        self.ckpt = None
        self.warmed_up = False
        self.restore_request: Optional[int] = None
        self.step_log: List[float] = []
        self.next_seq = 1                                  # ExecutePlan ordering when the client pipelines steps
        self.seq_cv = threading.Condition(self.exec_lock)
        from .. import _C
        self.env = _C.ServiceEnv.instance()
        self.env.load()

    # ---------------------------------------------------------------- master -> workers control plane
    def _bcast(self, msg: Optional[Dict[str, Any]]) -> Dict[str, Any]:
        if self.world == 1:
            return msg
        box = [msg]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def worker_loop(self) -> None:
        """Non-master ranks: execute whatever the master dispatches (DispatchPlan / ExecuteRemotePlan / save ...)."""
        while True:
            msg = self._bcast(None)
            if msg["cmd"] == "shutdown":
                return
            getattr(self, "_do_" + msg["cmd"])(msg)

    # ---------------------------------------------------------------- commands (run on all ranks)
    def _do_build(self, msg) -> int:
        from ..api import Trainer
        graph = Graph.from_dict(msg["graph"])
        tr = Trainer(graph, strategy=msg.get("strategy", self.strategy), device=self.device,
                     use_cuda_graph=self.use_cuda_graph, comm_mode=self.comm_mode, seed=msg.get("seed", 0))
        handle = self.cache.insert(tr)
        from ..ckpt import CheckpointManager
        self.ckpt = CheckpointManager(self.ckpt_root, self.rank, self.world, max_to_keep=msg.get("max_to_keep", 5))
        return handle

    def _do_execute(self, msg):
        tr = self.cache.get(msg["handle"])
        feeds = msg["feeds"]
        if self.env.get_bool("FAKE_INPUT"):      # cache the first step's inputs and reuse them
            if self.fake_input_cache is None:
                self.fake_input_cache = feeds
            feeds = self.fake_input_cache
        if self.restore_request is not None:      # restore happens during warm-up of the next ExecutePlan
            self.ckpt.restore(tr.executor(), self.restore_request if self.restore_request >= 0 else None)
            self.restore_request = None
        loss = tr.step(feeds)
        if not self.warmed_up:
            self.warmed_up = True
            if self.ckpt is not None:
                self.ckpt.maybe_lazy_save(tr.executor())
        return loss, tr

    def _do_save(self, msg):
        tr = self.cache.get(msg["handle"]) if msg.get("handle") else None
        self.ckpt.max_to_keep = msg.get("max_to_keep", self.ckpt.max_to_keep)
        if self.ckpt.request_save(msg["global_step"], self.warmed_up):
            return self.ckpt.save(tr.executor(), msg["global_step"])
        return "lazy"

    @staticmethod
    def _wants_slots(tr, names) -> bool:
        """Optimizer slots (moments, Adafactor / SM3 statistics) are gathered only when a requested name is not a variable."""
        variables = set(tr.executor().store.names.values())
        return any(k not in variables for k in names)

    def _do_sync_state(self, msg):
        """Make fp32 master weights / moments whole on every rank (sharded-optimizer plans keep only the owned chunk fresh)."""
        tr = self.cache.get(msg["handle"])
        # whole variables on the master, whatever the plan did to them (ZeRO chunks, stored shards, pipeline stages)
        return tr.full_state_dict(moments=bool(msg.get("moments")), dst=0)

    def _do_restore(self, msg):
        self.restore_request = msg.get("global_step", -1)
        return "pending"

    # ---------------------------------------------------------------- gRPC handlers (master only)
    def BuildExecutionPlan(self, req: bytes, ctx) -> bytes:
        m = unpack(req)
        msg = {"cmd": "build", **m}
        with self.exec_lock:
            self._bcast(msg)
            handle = self._do_build(msg)
        tr = self.cache.get(handle)
        return pack({"handle": handle, "plan_info": {k: v for k, v in tr.plan_info.items() if k != "log"}})

    def TransferToServerHost(self, req: bytes, ctx) -> bytes:
        m = unpack(req)     # {"name", "tensor" | ("shape","dtype"), "variable": bool}
        if m.get("variable"):
            return pack({"ok": True, "note": "variables are initialised on the servers from their init specs"})
        with self.exec_lock:                  # (ExecutePlan snapshots the registered inputs under the same lock)
            self.host_inputs[m["name"]] = m["tensor"]
        return pack({"ok": True})

    def ExecutePlan(self, req: bytes, ctx) -> bytes:
        import time
        m = unpack(req)
        seq = m.get("seq")
        with self.seq_cv:                     # == exec_lock; every broadcast + command pair runs under it
            if seq is not None:               # pipelined clients (NUM_PARALLEL_RPC_STEPS): steps run in the order they were issued
                ok = self.seq_cv.wait_for(lambda: seq <= self.next_seq, timeout=600)
                if seq < self.next_seq:
                    raise RuntimeError(f"ExecutePlan seq {seq} arrived after step {self.next_seq - 1} had run (duplicate or reordered request)")
                if not ok:
                    # the predecessor never arrived (lost request): give up on it so that the steps queued behind this one do not
                    # each sit out their own timeout
                    self.next_seq = seq + 1
                    self.seq_cv.notify_all()
                    raise RuntimeError(f"ExecutePlan seq {seq}: step {seq - 1} never arrived within 600 s")
            feeds = m.get("feeds") or {k: self.host_inputs[k] for k in m.get("input_names", self.host_inputs)}
            try:
                t0 = time.time()
                msg = {"cmd": "execute", "handle": m["handle"], "feeds": feeds}
                self._bcast(msg)
                loss, tr = self._do_execute(msg)
                dt = (time.time() - t0) * 1e3
                self.step_log.append(dt)
                out = {"loss": loss, "duration_ms": dt}
                want = m.get("fetch_vars")
                if want or m.get("fetch_all"):
                    msg = {"cmd": "sync_state", "handle": m["handle"], "moments": self._wants_slots(tr, want or [])}
                    self._bcast(msg)
                    sd = self._do_sync_state(msg)
                    names = want or [k for k in sd if not k.endswith(("/m", "/v"))]
                    out["vars"] = {k: sd[k].cpu() for k in names if k in sd}
            finally:
                if seq is not None:
                    self.next_seq = seq + 1
                    self.seq_cv.notify_all()
        return pack(out)

    def FetchResourceVars(self, req: bytes, ctx) -> bytes:
        m = unpack(req)
        tr = self.cache.get(m["handle"])
        with self.exec_lock:
            msg = {"cmd": "sync_state", "handle": m["handle"], "moments": self._wants_slots(tr, m.get("names") or [])}
            self._bcast(msg)
            sd = self._do_sync_state(msg)
        names = m.get("names") or [k for k in sd if not k.endswith(("/m", "/v"))]
        return pack({k: sd[k].cpu() for k in names if k in sd})

    def DoRemoteSave(self, req: bytes, ctx) -> bytes:
        m = unpack(req)
        msg = {"cmd": "save", **m}
        with self.exec_lock:
            self._bcast(msg)
            r = self._do_save(msg)
        return pack({"result": r})

    def DoRemoteRestore(self, req: bytes, ctx) -> bytes:
        m = unpack(req)
        msg = {"cmd": "restore", **m}
        with self.exec_lock:
            self._bcast(msg)
            r = self._do_restore(msg)
        return pack({"result": r})

    def GetServerInfo(self, req: bytes, ctx) -> bytes:
        return pack({"world": self.world, "config": self.env.dump(), "steps": len(self.step_log)})

    def Shutdown(self, req: bytes, ctx) -> bytes:
        with self.exec_lock:
            self._bcast({"cmd": "shutdown"})
        threading.Timer(0.2, lambda: self._server.stop(0)).start()
        return pack({"ok": True})


def serve(ip: str = "127.0.0.1", port: int = 0, block: bool = True, **kw):
    """Start the server job's endpoint.  Rank 0 serves gRPC; other ranks enter the worker loop."""
    impl = ServiceImpl(**kw)
    if impl.rank != 0:
        impl.worker_loop()
        return impl, None, None
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=8), options=_MAX)
    handlers = {m: grpc.unary_unary_rpc_method_handler(getattr(impl, m)) for m in METHODS}
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))
    bound = server.add_insecure_port(f"{ip}:{port}")
    impl._server = server
    server.start()
    if block:
        server.wait_for_termination()
    return impl, server, bound
