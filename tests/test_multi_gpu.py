"""The multi-GPU path on real devices: runs tests/multi_gpu/check_sharded.py under torchrun with every visible GPU (2..8).
Skipped on a single-GPU box; the gloo tests in test_sharded_cpu.py cover the host logic there."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_parity_on_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "multi-gpu parity ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
