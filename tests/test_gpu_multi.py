"""GPU, two or more devices: the multi-GPU paths with REAL NCCL / peer memory, one process per GPU (tests/multi/worker.py under
torch.distributed.run) -- the four-step transform against the oracle at BASELINE's 2^22, the point-split MSM, the round's
commitments dealt to the ranks, and the split proof against the single-GPU proof.  Skipped on a one-GPU box (the emulated-rank
tests in test_gpu_dist_ntt.py and the gloo tests in test_dist_gloo.py still run there)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_paths_against_oracle_and_single_gpu(gpu, tmp_path):
    count = gpu.device_count()
    if count < 2:
        pytest.skip("needs at least two GPUs (run under gpurun --gpus 2)")
    world = 2 if count < 4 else 4
    out = tmp_path / "multi.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi", "worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    checks = json.load(open(out))
    assert checks.pop("world") == world
    assert len(checks) >= 9 and all(checks.values()), checks
