"""CPU-only checks of bench.py's host logic: the reference arm runs and prints a well-formed
JSON line; under a 2-process gloo launch only rank 0 prints and the others exit 0."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_line(line):
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
              "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "objects/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
    return d


def test_reference_arm_single():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check_line(lines[0])


def test_reference_arm_two_ranks_rank0_only():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
        capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = _check_line(lines[0])
    assert d["n_gpus"] == 2


def test_gloo_max_over_ranks():
    """The N>1 timing reduction (max over ranks) on the gloo backend, world_size 2."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import bench
dist.init_process_group("gloo")
r = dist.get_rank()
v = bench.max_over_ranks(10.0 + r, 2, torch.device("cpu"))
bench.barrier(2)
assert v == 11.0, v
if r == 0: print("MAXOK")
dist.destroy_process_group()
''' % ROOT
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29534", "-c", code] if False else
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "_gloo_worker.py")],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAXOK" in out.stdout


def test_gloo_trainer_buckets_and_allreduce():
    """Data-parallel Trainer host logic on gloo, world_size 2 (tests/_gloo_trainer_worker.py)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29541",
         os.path.join(ROOT, "tests", "_gloo_trainer_worker.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "TRAINEROK" in out.stdout
