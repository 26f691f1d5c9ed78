"""Multi-GPU parity of the ring op (skipped unless >= 2 GPUs are visible): every rank's shard of
out/dq/dk/dv from the N-GPU ring equals the single-GPU result of the same kernels."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ring_on_all_visible_gpus():
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(ROOT, "tests", "ring_multi_gpu_worker.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "RING_MULTI_GPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
