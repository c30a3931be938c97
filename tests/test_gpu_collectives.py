"""Multi-GPU numerics of every collective kernel variant (one-shot / two-shot / NVLS / trees /
relay subsets / zero-copy) against fp32 references: spawns tests/gpu_collectives_worker.py under
torchrun on min(#GPUs, 4) ranks. Skipped on boxes with a single GPU (the single-GPU kernel paths are
covered by tests/test_gpu_ops.py)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import free_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_collectives_multi_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, ADAPCC_TIMEOUT_MS="15000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "gpu_collectives_worker.py"), "--quick"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "failures: 0" in r.stdout, tail


def _torchrun(worker, world, timeout=600, *args):
    env = dict(os.environ, ADAPCC_TIMEOUT_MS="15000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", worker), *args]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    return r, (r.stdout + r.stderr)[-4000:]


def test_engine_gradients_are_the_average_of_the_local_gradients():
    """The flat engine's bucketed all-reduce (sinks + hooks + side stream, eager and graph) against NCCL."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r, tail = _torchrun("gpu_engine_parity_worker.py", 4 if n >= 4 else 2)
    assert r.returncode == 0, tail
    assert "wrong averaged gradients: 0 of" in r.stdout, tail


def test_flag_protocol_soak():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r, tail = _torchrun("gpu_soak_worker.py", 4 if n >= 4 else 2, 600, "--ops", "3000")
    assert r.returncode == 0 and "failures: 0" in r.stdout, tail


def test_zero1_engine_matches_data_parallel():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, ADAPCC_TIMEOUT_MS="15000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "gpu_zero1_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "failures: 0" in r.stdout, tail
