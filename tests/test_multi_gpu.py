"""Sharded population on real GPUs (-m gpu, needs >= 2 devices; skipped on a one-GPU box): tools/multi_gpu_check.py under
torchrun -- every exchange mode (NCCL all-gather, NVLink peer gather, delta push) must reproduce the single-GPU run of the
same chains bit for bit, with outlier resets across shards, and the sharded R_stat must agree to 1e-9."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_run_equals_single_gpu_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (run tools/multi_gpu_check.py under torchrun on a multi-GPU box)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MULTI_GPU_CHECK PASS world={world}" in r.stdout
