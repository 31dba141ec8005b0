"""NCCL test of the multi-GPU path (skipped on a box with fewer than 2 GPUs): torchrun with 2
ranks, shards scanned per GPU, offsets gathered over NCCL, compared with the single-GPU result."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_scan_and_offset_gather():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    world = 4 if ngpu >= 4 else 2
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("NCCL_GATHER ")]
    assert line, r.stdout[-2000:]
    rep = json.loads(line[-1][len("NCCL_GATHER "):])
    assert rep["ip"]["ok"] and rep["lit64"]["ok"], rep
    assert rep["ip"]["matches"] > 0 and rep["lit64"]["matches"] > 0
