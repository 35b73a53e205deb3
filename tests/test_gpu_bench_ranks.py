"""GPU: bench.py launched the way the driver launches it for N>1 (torch.distributed.run, one process per rank, no extra
flags that skip legs).  Regression test for round 1's deadlock: ranks != 0 left before rank 0's train-step all-reduce.

A 1-GPU box cannot host two NCCL ranks (duplicate device), so there both ranks share cuda:0 over gloo
(`--backend gloo --single-device`, test hooks of bench.py); with >= 2 GPUs the exact driver command (NCCL) runs.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]      # rank 0 prints ONE line, the other rank none
    return json.loads(lines[0])


def test_two_rank_bench_with_every_leg(cuda):
    nccl = torch.cuda.device_count() >= 2
    line = _run([] if nccl else ["--backend", "gloo", "--single-device"])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["cpu_baseline"] is None                        # rank 0 at N=1 only
    t = line["train_step"]
    assert t["n_gpus"] == 2 and t["global_batch"] == 64 and len(t["ms_per_step_per_rank"]) == 2
    assert t["allreduce_bytes"] > 1_000_000 and t["allreduce_ms"] > 0


def test_reference_arm_under_torchrun_prints_one_line(cuda):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--cpu-patches", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert lines[0]["config"] == bench.bench_config(2)         # identical to the product arm's config
