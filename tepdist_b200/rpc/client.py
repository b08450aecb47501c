"""Client side of the RPC layer: what the user's training script talks to.

Reference parity (SURVEY F1/F2 + Appendix F): xla::Client::{BuildExecutionPlan, ExecutePlan, TransferToServerHost,
DoRemoteSave, DoRemoteRestore}; per-step input transfer; optional step pipelining NUM_PARALLEL_RPC_STEPS in [0,4]
(semaphore + background threads); variables fetched back only every FETCH_RESOURCE_VAR_STEPS; env SERVER_IP / SERVER_PORT.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional

import grpc
import torch

from ..ir import Graph
from .service import METHODS, SERVICE, _MAX, pack, unpack


class Client:
    def __init__(self, addr: Optional[str] = None):
        if addr is None:
            addr = f"{os.environ.get('SERVER_IP', '127.0.0.1')}:{os.environ.get('SERVER_PORT', '2222')}"
        self.channel = grpc.insecure_channel(addr, options=_MAX)
        self._stubs = {m: self.channel.unary_unary(f"/{SERVICE}/{m}") for m in METHODS}
        self.parallel_steps = min(4, max(0, int(os.environ.get("NUM_PARALLEL_RPC_STEPS", "0"))))
        self.fetch_every = int(os.environ.get("FETCH_RESOURCE_VAR_STEPS", "100000"))
        self._pool = ThreadPoolExecutor(max_workers=max(1, self.parallel_steps)) if self.parallel_steps else None
        self._sem = threading.Semaphore(max(1, self.parallel_steps))
        self._step = 0
        self.variables: Dict[str, torch.Tensor] = {}

    def _call(self, method: str, obj: Any) -> Any:
        return unpack(self._stubs[method](pack(obj)))

    def build_execution_plan(self, graph: Graph, strategy: Optional[str] = None, seed: int = 0, max_to_keep: int = 5) -> Dict[str, Any]:
        m = {"graph": graph.to_dict(), "seed": seed, "max_to_keep": max_to_keep}
        if strategy:
            m["strategy"] = strategy
        r = self._call("BuildExecutionPlan", m)
        self.handle = r["handle"]
        return r

    def transfer_to_server_host(self, name: str, tensor: Optional[torch.Tensor] = None, shape=None, dtype=None,
                                variable: bool = False) -> None:
        self._call("TransferToServerHost", {"name": name, "tensor": tensor, "shape": shape, "dtype": dtype, "variable": variable})

    def execute_plan(self, feeds: Optional[Dict[str, torch.Tensor]] = None, fetch_vars: Optional[List[str]] = None):
        """One training step.  With NUM_PARALLEL_RPC_STEPS > 0 returns a Future (bounded number of steps in flight)."""
        self._step += 1
        req = {"handle": self.handle, "feeds": feeds, "fetch_vars": fetch_vars, "seq": self._step}
        if fetch_vars is None and self.fetch_every and self._step % self.fetch_every == 0:
            req["fetch_all"] = True          # periodic refresh of every variable (FETCH_RESOURCE_VAR_STEPS)

        def run():
            try:
                r = self._call("ExecutePlan", req)
                if "vars" in r:
                    self.variables.update(r["vars"])
                return r
            finally:
                self._sem.release()

        self._sem.acquire()
        if self._pool is None:
            return run()
        return self._pool.submit(run)

    def fetch_resource_vars(self, names: Optional[List[str]] = None) -> Dict[str, torch.Tensor]:
        r = self._call("FetchResourceVars", {"handle": self.handle, "names": names})
        self.variables.update(r)
        return r

    def do_remote_save(self, global_step: int, max_to_keep: int = 5):
        return self._call("DoRemoteSave", {"handle": self.handle, "global_step": global_step, "max_to_keep": max_to_keep})

    def do_remote_restore(self, global_step: Optional[int] = None):
        return self._call("DoRemoteRestore", {"handle": self.handle, "global_step": -1 if global_step is None else global_step})

    def server_info(self):
        return self._call("GetServerInfo", {})

    def shutdown(self):
        try:
            return self._call("Shutdown", {})
        except grpc.RpcError:
            return None
