"""Control-plane workflow on CPU/gloo: DETECT -> PROFILE -> SYNTHESIS -> SETUP -> all_reduce ->
reconstruct_topology -> DDP hook with relay negotiation (coordinator over real gRPC), launcher CLI."""
import os
import socket
import sys
import tempfile

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_ports(n):
    """n DISTINCT free TCP ports (all sockets are held open until every port is known: two back-to-back single probes
    can return the same port)."""
    socks = [socket.socket() for _ in range(n)]
    try:
        for s in socks:
            s.bind(("127.0.0.1", 0))
        return [s.getsockname()[1] for s in socks]
    finally:
        for s in socks:
            s.close()


def _free_port():
    return _free_ports(1)[0]


def _worker(rank, world, port, coord_port, tmp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from types import SimpleNamespace

        from adapcc_b200 import ALLREDUCE
        from adapcc_b200.adapcc import AdapCC
        from adapcc_b200.strategy import Strategy

        args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "strategy", "auto.xml"),
                               logical_graph=os.path.join(tmp, "topology", "logical_graph.xml"), entry_point=6,
                               parallel_degree=2, profile_freq=2, backend="gloo", work_dir=tmp,
                               coordinator_port=coord_port, relay_threshold=0.05)
        os.makedirs(os.path.join(tmp, "strategy"), exist_ok=True)
        AdapCC.init(args, rank, rank, world)          # detect + profile + synthesise
        AdapCC.setup(ALLREDUCE)
        ok = True
        if rank == 0:
            ok &= os.path.exists(os.path.join(tmp, "topology", "topo_detect_0.xml"))
            ok &= os.path.exists(args.logical_graph) and os.path.exists(args.strategy_file)
            ok &= all(os.path.exists(os.path.join(tmp, "topology", f"topo_profile_{r}")) for r in range(world))
            st = Strategy.from_file(args.strategy_file)
            st.validate(world)
            # the synthesizer's per-message algorithm plan travels inside the strategy XML ...
            from adapcc_b200.synth.plan import AlgoPlan
            plan = AlgoPlan.from_attrs(st.attrs)
            ok &= plan is not None and len(plan.bands) >= 1 and plan.bands[-1][0] >= 1 << 60
            # ... and its thresholds in tunables.json
            import json as _json
            tun = _json.load(open(os.path.join(tmp, "topology", "tunables.json")))
            ok &= "bands" in tun and "one_shot_max_bytes" in tun and "ll_max_bytes" in tun
        comm = AdapCC.communicator
        ok &= getattr(comm, "plan", None) is not None                 # every rank loaded the same bands
        ok &= comm._resolve_algo(1 << 20, torch.float32, list(range(world))) in ("auto", "tree", "one_shot", "two_shot", "nvls", "ll")
        t = torch.full((1000,), float(rank + 1))
        comm.all_reduce(t, 1000)
        ok &= bool(torch.allclose(t, torch.full((1000,), world * (world + 1) / 2)))
        a2a = AdapCC.alltoall(torch.arange(world * 3, dtype=torch.float32) + 100 * rank)
        want = torch.cat([torch.arange(rank * 3, rank * 3 + 3, dtype=torch.float32) + 100 * src for src in range(world)])
        ok &= bool(torch.equal(a2a, want))
        # reduce-scatter / all-gather (ids 5 / 3: declared but never implemented by the reference): in place, shard layout
        # of the direct kernels; reduce-scatter followed by all-gather is an all-reduce
        from adapcc_b200 import ALLGATHER, REDUCESCATTER
        AdapCC.setup(REDUCESCATTER)
        AdapCC.setup(ALLGATHER)
        n = 1003                                                     # odd tail: the last shard is short
        x = torch.arange(n, dtype=torch.float32) * (rank + 1)
        lo, hi = AdapCC.reducescatter(x)
        full = torch.arange(n, dtype=torch.float32) * (world * (world + 1) / 2)
        ok &= 0 <= lo < hi <= n and bool(torch.equal(x[lo:hi], full[lo:hi]))
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        ok &= sorted(spans)[0][0] == 0 and sorted(spans)[-1][1] == n and all(a[1] == b[0] for a, b in zip(sorted(spans), sorted(spans)[1:]))
        AdapCC.allgather(x)
        ok &= bool(torch.equal(x, full))
        y = torch.full((64,), float(rank))
        lo, hi = AdapCC.reducescatter(y, op="avg")
        ok &= bool(torch.allclose(y[lo:hi], torch.full((hi - lo,), (world - 1) / 2)))
        AdapCC.reconstruct_topology(args, ALLREDUCE)   # clear + init + setup again (re-entrant)
        comm = AdapCC.communicator
        t = torch.full((77,), 2.0)
        comm.all_reduce(t, 77, op="avg")
        ok &= bool(torch.allclose(t, torch.full((77,), 2.0)))
        # DDP + hook with relay negotiation
        model = torch.nn.Linear(16, 4)
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        ddp.register_comm_hook(state=None, hook=comm.cuda_allreduce_hook)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
        for step in range(3):
            comm.update_relay(step)
            loss = ddp(torch.randn(8, 16)).pow(2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
        w = [torch.zeros_like(model.weight) for _ in range(world)]
        dist.all_gather(w, model.weight.detach())
        ok &= all(torch.allclose(w[0], x, atol=1e-6) for x in w)        # replicas stayed in sync
        ok &= len(comm.stats["hook_rpc_s"]) == 3
        # the training-loop pattern of the reference's ViT script: reconstruct_topology in the middle of training; DDP
        # still holds the hook of the communicator that is now cleared -> it must forward to the live one
        stale = comm
        AdapCC.reconstruct_topology(args, ALLREDUCE)
        live = AdapCC.communicator
        ok &= live is not stale and stale._live() is live
        for step in range(2):
            live.update_relay(step)
            loss = ddp(torch.randn(8, 16)).pow(2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
        dist.all_gather(w, model.weight.detach())
        ok &= all(torch.allclose(w[0], x, atol=1e-6) for x in w)
        ok &= len(live.stats["hook_rpc_s"]) == 2 and len(stale.stats["hook_rpc_s"]) == 3
        AdapCC.clear(ALLREDUCE)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_detect_profile_synth_setup_reconstruct_hook_cpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        port, cport = _free_ports(2)
        procs = [ctx.Process(target=_worker, args=(r, world, port, cport, tmp, q)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(180) for p in procs]
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()), res


def test_launcher_cli_parity(tmp_path, monkeypatch):
    from adapcc_b200 import launcher

    monkeypatch.chdir(tmp_path)
    a = launcher.build_parser().parse_args(["--num-process", "8", "--ips", "10.0.0.1:4,10.0.0.2:4", "--master",
                                            "10.0.0.1", "--exec-file", "train_ddp.py", "--socket_port", "5000",
                                            "--entry_point", "7", "--strategy_file", "s.xml", "--logical_graph", "g.xml",
                                            "--parallel_degree", "2", "--profile_freq", "100"])
    assert launcher.ip_table(launcher.parse_hosts(a.ips)) == ["10.0.0.1"] * 4 + ["10.0.0.2"] * 4
    cmds = launcher.commands(a)
    assert [h for h, _ in cmds] == ["10.0.0.1", "10.0.0.2"]
    flat = " ".join(cmds[1][1])
    for flag in ("--port=5000", "--entry_point=7", "--strategy_file=s.xml", "--logical_graph=g.xml",
                 "--parallel_degree=2", "--profile_freq=100", "--node-rank=1", "--nnodes=2"):
        assert flag in flat
    assert launcher.main(["--num-process", "2", "--ips", "127.0.0.1:2", "--dry-run"]) == 0
    assert (tmp_path / "topology" / "ip_table.txt").read_text() == "127.0.0.1\n127.0.0.1\n"


def test_fault_injector(monkeypatch):
    from adapcc_b200.utils.fault import FaultInjector

    monkeypatch.setenv("ADAPCC_STRAGGLERS", "1,3")
    monkeypatch.setenv("ADAPCC_STRAGGLE_MS", "40")
    monkeypatch.setenv("ADAPCC_KILL", "2@7")
    inj = FaultInjector.from_env(3)
    assert inj.delay_s(0) == 0 and abs(inj.delay_s(2) - 0.04) < 1e-9
    assert FaultInjector.from_env(0).delay_s(5) == 0
    assert FaultInjector.from_env(2).kill_at == {2: 7}
    import time
    t0 = time.time()
    inj.before_backward(3)
    assert time.time() - t0 >= 0.035 and inj.log


def test_gns_and_checkpoint_helpers(tmp_path):
    from adapcc_b200.utils.checkpoint import State
    from adapcc_b200.utils.gns import compute_gns

    torch.manual_seed(0)
    G = torch.randn(1000)
    small = torch.stack([(G + torch.randn(1000) * 3).pow(2).sum() for _ in range(64)])
    big = (G + torch.randn(1000) * 3 / 8).pow(2).sum()            # 64x larger batch -> 8x less noise
    gns, g2, s = compute_gns(small, big, 1, 64)
    assert 0.3 * 9000 < float(s) < 3 * 9000 and 0.3 * 1000 < float(g2) < 3 * 1000
    m = torch.nn.Linear(4, 4)
    st = State(m, torch.optim.SGD(m.parameters(), lr=0.1), epoch=3, step=17)
    st.save(tmp_path / "ck.pt")
    m2 = torch.nn.Linear(4, 4)
    st2 = State(m2, torch.optim.SGD(m2.parameters(), lr=0.1))
    assert st2.load(tmp_path / "ck.pt") and st2.epoch == 3 and st2.step == 17
    assert torch.equal(m.weight, m2.weight)


def test_meters_and_metrics_sink(tmp_path, capsys):
    import json

    from adapcc_b200.utils.meters import AverageMeter, MetricsSink, ProgressMeter, busbw_gbs

    m = AverageMeter("step", ":.2f")
    for v in (1.0, 3.0):
        m.update(v)
    assert m.avg == 2.0 and str(m) == "step 3.00 (2.00)"
    line = ProgressMeter(100, [m], prefix="it ").display(7)
    assert "[  7/100]" in line and "step 3.00" in capsys.readouterr().out
    sink = MetricsSink(str(tmp_path / "m" / "metrics.jsonl"), rank=3)
    sink.emit("allreduce", bytes=1 << 20, seconds=1e-3)
    rec = json.loads((tmp_path / "m" / "metrics.jsonl").read_text().strip())
    assert rec["rank"] == 3 and rec["event"] == "allreduce" and rec["bytes"] == 1 << 20
    MetricsSink(None).emit("noop")                              # disabled sink: silently ignored
    # nccl-tests factors: 8 ranks, 1 GB in 1 ms
    assert abs(busbw_gbs(10 ** 9, 1e-3, 8) - 1750.0) < 1e-6
    assert abs(busbw_gbs(10 ** 9, 1e-3, 8, "alltoall") - 875.0) < 1e-6
    assert abs(busbw_gbs(10 ** 9, 1e-3, 8, "boardcast") - 1000.0) < 1e-6


def test_elastic_example_checkpoints_and_resumes_on_cpu(tmp_path):
    """examples/elastic_imagenet.py (the reference's torchelastic workload) end to end on 2 gloo ranks: train one epoch,
    write the checkpoint, restart with a larger epoch budget and resume after the saved epoch."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pytest.importorskip("torchvision")

    def run(epochs):
        port, cport = _free_ports(2)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.join(root, "examples", "elastic_imagenet.py"),
               "--backend", "gloo", "--epochs", str(epochs), "--steps_per_epoch", "1", "--batch", "2",
               "--checkpoint", str(tmp_path / "ckpt.pt")]
        r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, PYTHONPATH=root, ADAPCC_COORD_PORT=str(cport)))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        return r.stdout

    out = run(1)
    assert "resuming from epoch 0" in out and (tmp_path / "ckpt.pt").exists()
    out = run(2)
    assert "resuming from epoch 1" in out and "Epoch: [1]" in out and "Epoch: [0]" not in out


def test_train_ddp_template_full_workflow_from_an_empty_directory(tmp_path):
    """The reference's template script (train_ddp.py) under torchrun on 2 gloo ranks with entry_point 6, started in an
    empty working directory: detect -> profile -> synthesise must create ./topology and ./strategy themselves."""
    import subprocess

    mport, cport = _free_ports(2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(mport), os.path.join(ROOT, "train_ddp.py"), "--backend", "gloo", "--model", "mlp",
           "--batch", "8", "--steps", "3", "--entry_point", "6"]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PYTHONPATH=ROOT, ADAPCC_COORD_PORT=str(cport)))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "step 2" in r.stdout
    assert (tmp_path / "strategy" / "strategy.xml").exists()
    for f in ("ip_table.txt", "logical_graph.xml", "topo_profile_0", "topo_profile_1", "tunables.json"):
        assert (tmp_path / "topology" / f).exists(), f


def test_wait_time_measurement_script_on_cpu(tmp_path):
    """adapcc_b200/bench/wait_time.py (the reference's units-test/wait-time measurement) on 2 gloo ranks: writes the
    per-step first-bucket gap CSV and prints the summary."""
    import subprocess

    mport, cport = _free_ports(2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(mport), "-m", "adapcc_b200.bench.wait_time", "--backend", "gloo", "--steps", "5",
           "--heter_alpha", "1.5"]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PYTHONPATH=ROOT, ADAPCC_COORD_PORT=str(cport)))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "mean" in r.stdout and "median" in r.stdout
    rows = (tmp_path / "wait_time.csv").read_text().strip().splitlines()
    assert len(rows) >= 3


def test_net_probe_bandwidth_and_latency_on_loopback():
    """adapcc_b200/bench/net_probe.py (the reference's iperf / ping cloud traces, cloud/band_profile.py) against its own
    server on loopback."""
    import threading

    from adapcc_b200.bench import net_probe

    with socket.socket() as s:                                 # a free port
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    threading.Thread(target=net_probe.serve, args=(port,), daemon=True).start()
    for _ in range(50):
        try:
            socket.create_connection(("127.0.0.1", port), timeout=0.2).close()
            break
        except OSError:
            import time
            time.sleep(0.05)
    assert net_probe.bandwidth("127.0.0.1", port, duration=0.2) > 0.1          # Gb/s
    lat = net_probe.latency("127.0.0.1", port, n=20)
    assert 0 < lat < 50                                                         # ms


def test_primitive_benchmark_main_on_cpu(tmp_path):
    """``python -m adapcc_b200.adapcc`` — the reference's benchmark ``__main__`` (adapcc.py:81-117): ones(16) * i through
    all_reduce / reduce / boardcast on 2 gloo ranks."""
    import subprocess

    mport, cport = _free_ports(2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(mport), "-m", "adapcc_b200.adapcc", "--backend", "gloo"]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PYTHONPATH=ROOT, ADAPCC_COORD_PORT=str(cport)))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("rank ")]
    assert len(lines) == 12
    assert "rank 0 allreduce: " + str([2.0] * 16) in lines and "rank 1 allreduce: " + str([4.0] * 16) in lines
    # reduce: slice t of the sum lands on tree t's root only (reference semantics); broadcast: the roots' data everywhere
    assert "rank 0 reduce: " + str([2.0] * 8 + [1.0] * 8) in lines and "rank 1 reduce: " + str([1.0] * 8 + [2.0] * 8) in lines
    assert "rank 1 boardcast: " + str([1.0] * 16) in lines


def test_training_survives_a_dead_worker_on_cpu(tmp_path):
    """Fault tolerance end to end (reference README: "continued communication without being blocked by the straggler /
    faulty"): 3 gloo ranks train through the DDP hook, rank 2 dies at step 2, ranks 0 and 1 finish all six steps with the
    active set shrunk to [0, 1]."""
    import subprocess

    mport, cport = _free_ports(2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr",
           "127.0.0.1", "--master-port", str(mport), os.path.join(ROOT, "tests", "cpu_fault_worker.py"), str(tmp_path),
           str(cport)]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PYTHONPATH=ROOT, ADAPCC_COORD_PORT=str(cport)))
    out = r.stdout + r.stderr
    assert "[rank 2] dying at step 2" in out
    for rank in (0, 1):
        assert f"[rank {rank}] finished" in out, out[-3000:]
        assert f"[rank {rank}] step 5 active [0, 1]" in out, out[-3000:]
    assert "[rank 0] step 1 active [0, 1, 2]" in out


def test_api_misuse_is_handled(tmp_path):
    """Calls in the wrong order must end in a clear error or be harmless — never in a leaked server or a hang."""
    import subprocess

    mport, cport = _free_ports(2)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpu_api_misuse_worker.py"), str(mport), str(cport)],
                       cwd=tmp_path, capture_output=True, text=True, timeout=180, env=dict(os.environ, PYTHONPATH=ROOT))
    out = r.stdout
    assert r.returncode == 0, (out + r.stderr)[-2000:]
    assert "use before init -> RuntimeError AdapCC.init(" in out
    assert "unknown prim -> NotImplementedError" in out
    for case in ("allreduce before setup", "double init", "setup twice", "setup other prims", "reduce", "clear unknown prim",
                 "clear", "clear twice", "use after clear"):
        assert f"{case} -> ok" in out, out


def test_hook_training_matches_stock_ddp_on_cpu(tmp_path):
    """Accuracy / precision check (the reference's accuracy benchmark): 30 SGD steps through the AdapCC comm hook end in
    the same parameters (to fp32 rounding) and the same accuracy as stock DDP gradient averaging — 3 gloo ranks."""
    import re
    import subprocess

    mport, cport = _free_ports(2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr",
           "127.0.0.1", "--master-port", str(mport), os.path.join(ROOT, "tests", "cpu_convergence_worker.py"), str(tmp_path),
           str(cport)]
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    rows = re.findall(r"\[rank (\d)\] max param diff ([0-9.e+-]+) acc stock ([0-9.]+) acc hook ([0-9.]+)", r.stdout)
    assert len(rows) == 3, r.stdout
    for _, diff, a0, a1 in rows:
        assert float(diff) < 1e-5 and float(a0) > 0.9 and abs(float(a0) - float(a1)) < 1e-6


def test_synthetic_ddp_benchmark_on_cpu(tmp_path):
    """adapcc_b200/bench/synthetic_ddp.py (the reference's Horovod synthetic img/s benchmark, nccl-perf/pytorch_synthetic.py)
    on 2 gloo ranks through the communicator's hook, with the Horovod script's ``--fp16-allreduce`` switch."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "-m", "adapcc_b200.bench.synthetic_ddp", "--backend", "gloo", "--model", "resnet18",
           "--image_size", "32", "--num-classes", "10", "--batch-size", "8", "--num-warmup-batches", "1",
           "--num-batches-per-iter", "2", "--num-iters", "2", "--fp16-allreduce"]
    import subprocess
    r = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-2500:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("Img/sec per GPU:")]
    assert len(line) == 1 and "total on 2 GPU(s)" in line[0] and "wire=float16" in line[0]
    assert float(line[0].split(":")[1].split()[0]) > 0


def test_doctor_reports_environment_and_rpc_latency():
    """python -m adapcc_b200.doctor (pre-flight check: the reference's check_mpi_connect / RPC latency dumps): toolchain,
    native library, GPUs, a 2-process rendezvous on loopback and the coordinator's gRPC round trip."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "adapcc_b200.doctor", "--ranks", "2", "--rpc", "100", "--json"], cwd=root,
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    rep = json.loads(r.stdout[r.stdout.index("{"):])
    assert rep["library"]["built"] and rep["library"]["missing_symbols"] == [] and rep["problems"] == []
    assert rep["rendezvous"]["ok"] and rep["rendezvous"]["backend"] == "gloo" and rep["rendezvous"]["ranks"] == 2
    c = rep["coordinator_rpc"]
    assert c["calls"] == 100 and 0 < c["median_ms"] <= c["p95_ms"] <= c["max_ms"] and c["median_ms"] < 50
