"""CPU tests: client -> gRPC -> service -> planner -> runtime round trip, checkpoint save / rotate / lazy / restore."""
import json
import os

import pytest

import torch

from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph
from tepdist_b200.rpc.client import Client
from tepdist_b200.rpc.service import serve


def _start(tmp_path, **kw):
    impl, server, port = serve("127.0.0.1", 0, block=False, device=torch.device("cpu"), ckpt_root=str(tmp_path), **kw)
    return impl, server, Client(f"127.0.0.1:{port}")


def test_client_server_training_and_fetch(tmp_path):
    impl, server, cl = _start(tmp_path)
    cfg = CONFIGS["tiny"]
    g = build_gpt2_graph(cfg)
    r = cl.build_execution_plan(g)
    assert r["handle"] >= 1
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    lab = torch.roll(tok, -1, 1)
    cl.transfer_to_server_host("tokens", tok)
    cl.transfer_to_server_host("labels", lab)
    cl.transfer_to_server_host("model/wte", shape=[cfg.padded_vocab, cfg.n_embd], dtype="bf16", variable=True)   # shape only
    losses = [cl.execute_plan()["loss"] for _ in range(3)]           # inputs registered on the server host
    losses += [cl.execute_plan({"tokens": tok, "labels": lab})["loss"]]
    assert losses[-1] < losses[0]
    v = cl.fetch_resource_vars(["model/ln_f/g"])
    assert v["model/ln_f/g"].shape == (cfg.n_embd,)
    out = cl.execute_plan({"tokens": tok, "labels": lab}, fetch_vars=["model/ln_f/b"])
    assert "model/ln_f/b" in out["vars"] and out["duration_ms"] > 0
    info = cl.server_info()
    assert info["world"] == 1 and "OPT_LEVEL" in info["config"] and info["steps"] == 5
    server.stop(0)


def test_optimizer_settings_travel_over_rpc(tmp_path):
    """A graph with Adafactor in relative-step mode (lr = None), global-norm clipping and a schedule spec -- None values and nested
    dicts in graph.meta -- is serialised to the server, trains there like it does locally, and its reduced-shape slots can be fetched."""
    from tepdist_b200.runtime.executor import Executor
    impl, server, cl = _start(tmp_path)
    cfg = CONFIGS["tiny"]
    g = build_gpt2_graph(cfg, optimizer="adafactor", clip_norm="global", clip_norm_value=0.5,
                         schedule={"kind": "warmup_linear_decay", "warmup_steps": 2, "total_steps": 6})
    g.meta["optimizer"]["lr"] = None
    assert cl.build_execution_plan(g)["handle"] >= 1
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
    remote = [cl.execute_plan(feeds)["loss"] for _ in range(3)]
    local = Executor(g, torch.device("cpu"), use_cuda_graph=False)
    for a in remote:
        assert a == pytest.approx(float(local.step(feeds)[0]), rel=1e-6)
    v = cl.fetch_resource_vars(["model/h0/attn/c_attn/w/vr"])
    assert tuple(v["model/h0/attn/c_attn/w/vr"].shape) == (3 * cfg.n_embd,)
    server.stop(0)


def test_checkpoint_lazy_save_rotate_restore(tmp_path):
    impl, server, cl = _start(tmp_path)
    cfg = CONFIGS["tiny"]
    cl.build_execution_plan(build_gpt2_graph(cfg), max_to_keep=2)
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
    assert cl.do_remote_save(0)["result"] == "lazy"           # before the first step: deferred
    l0 = cl.execute_plan(feeds)["loss"]                       # warm-up step performs the lazy save
    ck = os.path.join(str(tmp_path), "ckpt_0_of_1")
    assert os.path.isdir(os.path.join(ck, "step_0"))
    for s in (1, 2, 3):
        cl.execute_plan(feeds)
        cl.do_remote_save(s, max_to_keep=2)
    assert sorted(d for d in os.listdir(ck) if d.startswith("step_")) == ["step_2", "step_3"]      # rotation
    assert json.load(open(os.path.join(ck, "checkpoint_queue.json"))) == [2, 3]
    w_before = cl.fetch_resource_vars(["model/h0/mlp/c_fc/w"])["model/h0/mlp/c_fc/w"].clone()
    l_a = cl.execute_plan(feeds)["loss"]
    cl.do_remote_restore(3)                                   # takes effect at the next ExecutePlan
    l_b = cl.execute_plan(feeds)["loss"]
    assert abs(l_a - l_b) < 1e-6                              # same state => same loss as the step right after save 3
    manifest = json.load(open(os.path.join(ck, "step_3", "manifest.json")))
    assert manifest["vars"]["model/h0/mlp/c_fc/w"]["full_shape"] == [4 * cfg.n_embd, cfg.n_embd]
    server.stop(0)


def test_step_pipelining_env(tmp_path, monkeypatch):
    monkeypatch.setenv("NUM_PARALLEL_RPC_STEPS", "2")
    impl, server, cl = _start(tmp_path)
    cfg = CONFIGS["tiny"]
    cl.build_execution_plan(build_gpt2_graph(cfg))
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    futs = [cl.execute_plan({"tokens": tok, "labels": torch.roll(tok, -1, 1)}) for _ in range(4)]
    losses = [f.result()["loss"] for f in futs]
    assert len(losses) == 4 and losses[-1] < losses[0]
    server.stop(0)


def test_periodic_variable_fetch_and_step_order(tmp_path, monkeypatch):
    """FETCH_RESOURCE_VAR_STEPS=2: every second ExecutePlan returns all (non-slot) variables and refreshes Client.variables;
    pipelined steps (NUM_PARALLEL_RPC_STEPS) execute in issue order -- the server enforces the client's sequence numbers."""
    monkeypatch.setenv("FETCH_RESOURCE_VAR_STEPS", "2")
    monkeypatch.setenv("NUM_PARALLEL_RPC_STEPS", "3")
    impl, server, cl = _start(tmp_path)
    cfg = CONFIGS["tiny"]
    cl.build_execution_plan(build_gpt2_graph(cfg))
    tok = torch.randint(0, cfg.n_vocab, (cfg.batch, cfg.n_ctx), dtype=torch.int32)
    feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
    r1 = cl.execute_plan(feeds).result()
    assert "vars" not in r1 and not cl.variables
    r2 = cl.execute_plan(feeds).result()
    assert "model/ln_f/g" in r2["vars"] and not any(k.endswith(("/m", "/v")) for k in r2["vars"])
    snap = cl.variables["model/ln_f/g"].clone()
    futs = [cl.execute_plan(feeds) for _ in range(6)]
    losses = [f.result()["loss"] for f in futs]
    assert losses == sorted(losses, reverse=True), losses       # same batch every step: in-order execution is monotone
    assert not torch.equal(cl.variables["model/ln_f/g"], snap)  # refreshed at steps 4, 6, 8
    assert impl.next_seq == 9
    server.stop(0)


def test_requests_cannot_carry_code(tmp_path):
    """Request bodies go through the restricted unpickler: a pickle that would run code on load is rejected by the server."""
    import io
    import pickle
    import grpc
    from tepdist_b200.rpc.service import SERVICE, unpack

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    buf = io.BytesIO()
    torch.save({"x": Evil()}, buf)
    with pytest.raises(pickle.UnpicklingError):
        unpack(buf.getvalue())
    impl, server, cl = _start(tmp_path)
    with pytest.raises(grpc.RpcError):
        cl.channel.unary_unary(f"/{SERVICE}/FetchResourceVars")(buf.getvalue())
    assert not (tmp_path / "pwned").exists()
    server.stop(0)


def test_resaving_a_step_keeps_its_directory(tmp_path):
    from tepdist_b200.ckpt import CheckpointManager
    from tepdist_b200.runtime.executor import Executor
    ex = Executor(build_gpt2_graph(CONFIGS["tiny"]), torch.device("cpu"), use_cuda_graph=False)
    cm = CheckpointManager(str(tmp_path), 0, 1, max_to_keep=1)
    cm.save(ex, 5)
    cm.save(ex, 5)
    assert cm.queue == [5] and os.path.exists(os.path.join(cm.dir, "step_5", "manifest.json"))
    assert cm.restore(ex) == 5


def test_a_step_counts_only_once_its_save_was_committed(tmp_path):
    """Every rank publishes its shards, then (after a barrier over the job) writes a COMMITTED marker into its step directory.
    A step whose marker is missing -- a rank died between the two -- is skipped by latest() and refused by restore()."""
    from tepdist_b200.ckpt import CheckpointManager
    from tepdist_b200.runtime.executor import Executor
    ex = Executor(build_gpt2_graph(CONFIGS["tiny"]), torch.device("cpu"), use_cuda_graph=False)
    cm = CheckpointManager(str(tmp_path), 0, 1, max_to_keep=3)
    cm.save(ex, 1)
    cm.save(ex, 2)
    assert cm.latest() == 2 and cm.committed(2)
    os.remove(os.path.join(cm.dir, "step_2", CheckpointManager.COMMIT))       # "the job died while saving step 2"
    fresh = CheckpointManager(str(tmp_path), 0, 1, max_to_keep=3)             # a restarted job
    assert fresh.latest() == 1 and fresh.restore(ex) == 1
    with pytest.raises(FileNotFoundError, match="never committed"):
        fresh.restore(ex, 2)
    assert fresh._latest_any_layout() == 1                                    # (the cross-plan restore path picks the same step)


def test_launcher_cluster_spec(tmp_path):
    from tepdist_b200.launch import entry_for
    spec = {"master": {"ip": "10.0.0.1", "port": 2222, "gpu_ids": [0, 1]}, "workers": [{"ip": "10.0.0.2", "port": 2223, "gpu_ids": [2, 3]}]}
    assert entry_for(spec, 1)["port"] == 2223
    spec["workers"][0]["gpu_ids"] = [2]
    try:
        entry_for(spec, 0)
        assert False
    except ValueError:
        pass
    # the reference's own template form (tf_tepdist/config_*_template.json): strings for the port and the GPU list, "localhost"
    ref_form = {"master": {"ip": "localhost", "port": "2222", "gpu_ids": "1,2"}}
    e = entry_for(ref_form, 0)
    assert e == {"ip": "127.0.0.1", "port": 2222, "gpu_ids": [1, 2]}


def test_service_env_flags_drive_planner(monkeypatch, tmp_path):
    """ServiceEnv keys set in the environment / CONFIG_FILE reach the planner (reference service_env.h:46-74)."""
    import json as _json
    from tepdist_b200 import config
    from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph
    from tepdist_b200.parallel import classify_parallelism, plan_spmd
    g = build_gpt2_graph(CONFIGS["tiny"], batch=4)
    # defaults: nothing explicit -> data parallel
    for k in ("VAR_MEM_LIMIT", "RULE_MODE", "NUM_STAGES", "NUM_MICRO_BATCHES", "FP16_COMM", "CONFIG_FILE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.chdir(tmp_path)
    config.env(reload=True)
    assert config.spmd_overrides() == {} and config.resolve_strategy("auto") == "auto" and config.comm_dtype() is None
    _, info = plan_spmd(g, 2, "auto")
    assert classify_parallelism(info, 2).startswith("dp")
    # VAR_MEM_LIMIT from the environment forces weights to be stored sharded -> tensor parallel
    monkeypatch.setenv("VAR_MEM_LIMIT", "1")
    config.env(reload=True)
    assert config.spmd_overrides()["var_mem_limit"] == 1.0
    _, info = plan_spmd(g, 2, "auto")
    assert classify_parallelism(info, 2).startswith("tp")
    monkeypatch.delenv("VAR_MEM_LIMIT")
    # JSON config file: pipeline config mode + 16-bit communication; environment beats the file
    cfg = tmp_path / "cfg.json"
    cfg.write_text(_json.dumps({"NUM_STAGES": "2", "NUM_MICRO_BATCHES": "4", "FP16_COMM": "true", "RULE_MODE": "false"}))
    monkeypatch.setenv("CONFIG_FILE", str(cfg))
    config.env(reload=True)
    assert config.resolve_strategy("auto") == "pp2m4" and config.resolve_strategy("tp") == "tp"
    import torch
    assert config.comm_dtype() == torch.bfloat16
    monkeypatch.setenv("NUM_STAGES", "1")
    monkeypatch.setenv("RULE_MODE", "true")
    config.env(reload=True)
    assert config.resolve_strategy("auto") == "rule"
    monkeypatch.delenv("CONFIG_FILE"); monkeypatch.delenv("NUM_STAGES"); monkeypatch.delenv("RULE_MODE")
    config.env(reload=True)


def test_trainer_save_restore_resumes_identically(tmp_path):
    """Trainer.save / restore: resuming from a checkpoint reproduces the loss trajectory (weights + AdamW moments + step)."""
    import torch
    from tepdist_b200.api import Trainer
    from tepdist_b200.models.gpt2 import CONFIGS, build_gpt2_graph
    cfg = CONFIGS["tiny"]
    g = build_gpt2_graph(cfg, batch=2)
    torch.manual_seed(0)
    tok = torch.randint(0, cfg.n_vocab, (2, cfg.n_ctx), dtype=torch.int32)
    feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
    tr = Trainer(g, device=torch.device("cpu"), use_cuda_graph=False)
    for _ in range(2):
        tr.step(feeds)
    tr.save(str(tmp_path), global_step=2, max_to_keep=2)
    after = [tr.step(feeds) for _ in range(3)]
    tr2 = Trainer(g, device=torch.device("cpu"), use_cuda_graph=False, seed=123)     # different init: everything must come from disk
    assert tr2.restore(str(tmp_path)) == 2
    resumed = [tr2.step(feeds) for _ in range(3)]
    for a, b in zip(after, resumed):
        assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (after, resumed)
    for s in (3, 4, 5):
        tr.save(str(tmp_path), global_step=s, max_to_keep=2)
    kept = sorted(os.listdir(os.path.join(str(tmp_path), "ckpt_0_of_1")))
    assert [k for k in kept if k.startswith("step_")] == ["step_4", "step_5"], kept


@pytest.mark.parametrize("strategy", ["auto", "pp2m2"])
def test_two_rank_server_job_via_launcher(tmp_path, strategy):
    """launch.py starts a 2-process server job from a cluster spec (master rank serves gRPC, the other rank sits in the
    worker loop); a client builds a plan (the master plans and dispatches it), trains, saves a sharded checkpoint and shuts
    the job down.  Losses equal the single-process run (reference: ExecutionCoordinator + ExecuteRemotePlan, E2-E6)."""
    import socket
    import subprocess
    import sys
    import time
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    spec = tmp_path / "cluster.json"
    spec.write_text(json.dumps({"master": {"ip": "127.0.0.1", "port": port, "gpu_ids": [0, 1]}, "workers": []}))
    env = dict(os.environ, OMP_NUM_THREADS="2", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cwd = str(tmp_path)
    job = subprocess.Popen([sys.executable, "-m", "tepdist_b200.launch", "--cluster", str(spec), "--task-index", "0", "--platform", "cpu"],
                           env=env, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                socket.create_connection(("127.0.0.1", port), timeout=0.5).close()
                break
            except OSError:
                assert job.poll() is None, job.stdout.read()[-3000:]
                time.sleep(0.5)
        cl = Client(f"127.0.0.1:{port}")
        cfg = CONFIGS["tiny"]
        g = build_gpt2_graph(cfg, batch=4)
        r = cl.build_execution_plan(g, strategy=strategy)
        assert r["plan_info"]["world"] == 2 and r["plan_info"]["parallelism"].startswith("dp" if strategy == "auto" else "pp2"), r
        torch.manual_seed(0)
        tok = torch.randint(0, cfg.n_vocab, (4, cfg.n_ctx), dtype=torch.int32)
        feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
        losses = [cl.execute_plan(feeds)["loss"] for _ in range(3)]
        assert cl.server_info()["world"] == 2
        fetched = cl.fetch_resource_vars()          # whole variables, wherever the plan put their pieces (ZeRO chunks / stages)
        cl.do_remote_save(3)
        cl.shutdown()
        job.wait(timeout=60)
    finally:
        if job.poll() is None:
            job.kill()
    from tepdist_b200.runtime.executor import Executor
    ref = Executor(g, torch.device("cpu"), use_cuda_graph=False)
    for a in losses:
        b = float(ref.step(feeds)[0])
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (losses, b)
    assert os.path.isdir(tmp_path / "ckpt_0_of_2" / "step_3") and os.path.isdir(tmp_path / "ckpt_1_of_2" / "step_3")
    want = {k: v for k, v in ref.store.state_dict().items() if not k.endswith(("/m", "/v"))}
    assert set(fetched) == set(want), set(fetched) ^ set(want)
    for k, v in want.items():
        assert tuple(fetched[k].shape) == tuple(v.shape) and torch.allclose(fetched[k].float(), v, atol=2e-4), k


def test_every_flag_override_key_exists_on_its_cxx_options_struct(monkeypatch, tmp_path):
    """parallel/__init__.py applies config overrides with `if hasattr(o, k)`: a key that does not exist on the pybind options
    object would be dropped silently and the flag would do nothing.  Set every flag and check every emitted key lands."""
    from tepdist_b200 import _C, config
    monkeypatch.chdir(tmp_path)
    for k, v in {"VAR_MEM_LIMIT": "123", "COST_FACTOR": "2.0", "OPT_LEVEL": "1", "IGNORE_ANNOTATION": "false", "AUX_AFFINITY": "true",
                 "FORWARD_SUB_GRAPH_NUM": "3", "ILP_TIME_LIMIT": "2", "UNBALANCED_RATIO": "0.2", "RULE_MODE": "true",
                 "MICRO_NUM_LIMIT": "3", "BUFFER_SAVE": "false", "MULTI_REORDER": "false"}.items():
        monkeypatch.setenv(k, v)
    monkeypatch.delenv("CONFIG_FILE", raising=False)
    config.env(reload=True)
    try:
        for overrides, obj in ((config.spmd_overrides(), _C.SpmdOptions()), (config.auto_parallel_overrides(), _C.AutoParallelOptions()),
                               (config.schedule_overrides(), _C.ScheduleOptions())):
            assert overrides, type(obj).__name__
            for k, v in overrides.items():
                assert hasattr(obj, k), f"{type(obj).__name__} has no field '{k}': the flag would be ignored"
                setattr(obj, k, v)
                got = getattr(obj, k)
                assert got == v or abs(float(got) - float(v)) < 1e-9, (k, got, v)
        so = config.spmd_overrides()
        assert so["var_mem_limit"] == 123.0 and so["opt_level"] == 1 and so["ilp_time_limit_s"] == 120.0 and so["aux_affinity"] is True
    finally:
        monkeypatch.undo()
        config.env(reload=True)


def test_cluster_spec_with_a_worker_entry_forms_one_job(tmp_path):
    """Two launcher invocations (task index 0 = master entry, 1 = worker entry, one process each) must join ONE server job:
    the client sees world == 2 and trains with single-process parity."""
    import socket
    import subprocess
    import sys
    import time

    def free_port():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
        return p
    port, rdzv = free_port(), free_port()
    spec = tmp_path / "cluster.json"
    spec.write_text(json.dumps({"master": {"ip": "127.0.0.1", "port": port, "gpu_ids": [0]},
                                "workers": [{"ip": "127.0.0.1", "port": port + 1, "gpu_ids": [0]}], "rdzv_port": rdzv}))
    env = dict(os.environ, OMP_NUM_THREADS="2", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    jobs = [subprocess.Popen([sys.executable, "-m", "tepdist_b200.launch", "--cluster", str(spec), "--task-index", str(i), "--platform", "cpu"],
                             env=env, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for i in (0, 1)]
    try:
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                socket.create_connection(("127.0.0.1", port), timeout=0.5).close()
                break
            except OSError:
                for j in jobs:
                    assert j.poll() is None, j.stdout.read()[-3000:]
                time.sleep(0.5)
        cl = Client(f"127.0.0.1:{port}")
        cfg = CONFIGS["tiny"]
        g = build_gpt2_graph(cfg, batch=4)
        r = cl.build_execution_plan(g, strategy="auto")
        assert r["plan_info"]["world"] == 2, r
        torch.manual_seed(0)
        tok = torch.randint(0, cfg.n_vocab, (4, cfg.n_ctx), dtype=torch.int32)
        feeds = {"tokens": tok, "labels": torch.roll(tok, -1, 1)}
        losses = [cl.execute_plan(feeds)["loss"] for _ in range(2)]
        cl.shutdown()
        for j in jobs:
            j.wait(timeout=60)
    finally:
        for j in jobs:
            if j.poll() is None:
                j.kill()
    from tepdist_b200.runtime.executor import Executor
    ref = Executor(g, torch.device("cpu"), use_cuda_graph=False)
    for a in losses:
        b = float(ref.step(feeds)[0])
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (losses, b)


def test_newly_wired_flags_have_an_observable_effect_and_no_key_is_unaccounted(monkeypatch, tmp_path):
    import re
    from tepdist_b200 import config
    from tepdist_b200.parallel import plan_pipeline, plan_spmd
    monkeypatch.chdir(tmp_path)
    for k in ("HW_PROFILE", "PP_BANDWIDTH", "NUM_GRADIENTS", "MULTI_REORDER", "CONFIG_FILE"):
        monkeypatch.delenv(k, raising=False)
    config.env(reload=True)
    g = build_gpt2_graph(CONFIGS["tiny"], batch=8)
    try:
        # HW_PROFILE: the evaluator's estimate is computed with different constants
        base = plan_pipeline(g, 4, 2, 4)[1]
        assert config.hw_profile().name != "" and config.hw_profile().flops > 1e14
        monkeypatch.setenv("HW_PROFILE", "reference_v100")
        config.env(reload=True)
        assert config.hw_profile().flops < 1e14
        v100 = plan_pipeline(g, 4, 2, 4)[1]
        # (strictly larger, not "N x": for this tiny model fixed per-stage terms dominate the scheduled makespan)
        assert v100["makespan_est"] > base["makespan_est"], (base["makespan_est"], v100["makespan_est"])
        monkeypatch.delenv("HW_PROFILE")
        # PP_BANDWIDTH (GB/s): slower pipeline links stretch the scheduled makespan
        monkeypatch.setenv("PP_BANDWIDTH", "0.001")
        config.env(reload=True)
        slow = plan_pipeline(g, 4, 2, 4)[1]
        assert slow["makespan_est"] > base["makespan_est"], (base["makespan_est"], slow["makespan_est"])
        monkeypatch.delenv("PP_BANDWIDTH")
        # NUM_GRADIENTS: sanity check against the number of updated variables
        n_apply = sum(1 for n in g.nodes if n.op.startswith("apply_"))
        monkeypatch.setenv("NUM_GRADIENTS", str(n_apply))
        config.env(reload=True)
        plan_spmd(g, 2, "auto")
        monkeypatch.setenv("NUM_GRADIENTS", str(n_apply + 1))
        config.env(reload=True)
        with pytest.raises(ValueError):
            plan_spmd(g, 2, "auto")
        monkeypatch.delenv("NUM_GRADIENTS")
        # MULTI_REORDER reaches the scheduler options
        monkeypatch.setenv("MULTI_REORDER", "false")
        config.env(reload=True)
        assert config.schedule_overrides() == {"reorder_send": False}
        monkeypatch.delenv("MULTI_REORDER")
        # every ServiceEnv key is either consumed by config.py / rpc/service.py or listed as inert, nothing in between
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        keys = set(re.findall(r"X\((\w+),", open(os.path.join(root, "tepdist_b200", "csrc", "service_env.h")).read()))
        cfg_src = open(os.path.join(root, "tepdist_b200", "config.py")).read()
        cfg_src = re.sub(r"# Accepted for compatibility.*?INERT_KEYS = \([^)]*\)", "", cfg_src, flags=re.S)   # the inert list itself
        assert "INERT_KEYS" not in cfg_src
        used_src = cfg_src + open(os.path.join(root, "tepdist_b200", "rpc", "service.py")).read()
        consumed = {k for k in keys if re.search(r"""["']%s["']""" % k, used_src)}
        assert consumed | set(config.INERT_KEYS) == keys, (sorted(keys - consumed - set(config.INERT_KEYS)), sorted(set(config.INERT_KEYS) - keys))
        assert not (consumed & set(config.INERT_KEYS)), sorted(consumed & set(config.INERT_KEYS))
    finally:
        monkeypatch.undo()
        config.env(reload=True)
