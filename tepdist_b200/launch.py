"""Server launcher (reference tf_tepdist/launch_worker.sh + config_*worker_template.json).

    python -m tepdist_b200.launch --cluster cluster.json --task-index 0

cluster.json: {"master": {"ip": "...", "port": 2222, "gpu_ids": [0,1,2,3]}, "workers": [{"ip":..., "port":..., "gpu_ids": [...]}]}
Run it once per entry (task index 0 = master, 1.. = workers): the entry decides CUDA_VISIBLE_DEVICES, one process per listed
GPU is spawned through torch.distributed.run and ALL entries join one rendezvous hosted by the master entry, so the spec
describes a single server job; global rank 0 (on the master entry) serves the gRPC endpoint master.ip:master.port (all
entries must list the same number of GPUs, as in the reference).
Tested with two entries on localhost (CPU / gloo).  Entries on different machines are untested; the peer-memory kernels rely
on CUDA IPC and are intra-node only, so such a job would have to run with COMM_MODE=nccl.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys


def _normalise(spec: dict) -> dict:
    """The reference's templates write `"gpu_ids": "0,1,2,3"` and `"port": "2222"` as strings; accept both forms."""
    for e in [spec["master"]] + list(spec.get("workers", [])):
        if isinstance(e.get("gpu_ids"), str):
            e["gpu_ids"] = [int(x) for x in e["gpu_ids"].replace(" ", "").split(",") if x != ""]
        e["port"] = int(e["port"])
        if e.get("ip") == "localhost":
            e["ip"] = "127.0.0.1"      # (the container's hostname resolution is not reliable; the loopback address is)
    return spec


def entry_for(spec: dict, task_index: int) -> dict:
    _normalise(spec)
    entries = [spec["master"]] + list(spec.get("workers", []))
    counts = {len(e["gpu_ids"]) for e in entries}
    if len(counts) != 1:
        raise ValueError("all workers must have equal GPU counts")
    return entries[task_index]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cluster", required=True)
    ap.add_argument("--task-index", type=int, default=0)
    ap.add_argument("--platform", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--strategy", default="auto")
    ap.add_argument("--serve-rank", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--ip", default=None)
    ap.add_argument("--port", type=int, default=None)
    args = ap.parse_args(argv)
    spec = json.load(open(args.cluster))
    e = entry_for(spec, args.task_index)
    if args.serve_rank:   # inside torchrun: one process per GPU
        import torch
        from .rpc.service import serve
        dev = torch.device("cpu") if args.platform == "cpu" else None
        serve(args.ip or e["ip"], args.port or e["port"], block=True, strategy=args.strategy, device=dev)
        return 0
    env = dict(os.environ, CLUSTER_SPEC=os.path.abspath(args.cluster), FRONTEND="torch")
    if args.platform == "cuda":
        env["CUDA_VISIBLE_DEVICES"] = ",".join(str(g) for g in e["gpu_ids"])
        env.setdefault("NCCL_DEBUG", "WARN")
    else:
        env["CUDA_VISIBLE_DEVICES"] = ""
    n = len(e["gpu_ids"])
    entries = [spec["master"]] + list(spec.get("workers", []))
    # ONE job over all entries (reference: launch_worker.sh is run once per task index and the servers find each other through
    # the cluster spec): entry i is node i of a torchrun rendezvous hosted by the master entry; the gRPC endpoint is served
    # by global rank 0 only, i.e. on the master entry.  `rdzv_port` may be given in the spec (default: master port + 1000).
    master = spec["master"]
    rdzv_port = int(spec.get("rdzv_port", int(master["port"]) + 1000))
    cmd = [sys.executable, "-m", "torch.distributed.run", f"--nnodes={len(entries)}", f"--node-rank={args.task_index}",
           f"--nproc-per-node={n}", "--master-addr", str(master["ip"]), "--master-port", str(rdzv_port),
           "-m", "tepdist_b200.launch", "--cluster", args.cluster, "--task-index", str(args.task_index), "--platform", args.platform,
           "--strategy", args.strategy, "--serve-rank", "--ip", str(master["ip"]), "--port", str(master["port"])]
    return subprocess.call(cmd, env=env)


if __name__ == "__main__":
    sys.exit(main())
