"""N > 1 on real GPUs (skipped on a one-GPU box): the library's NCCL exchange -- device-side all-reduce of the
ensemble sums and gather of the forest records to rank 0 -- against the same forest grown on one GPU."""
import os
import socket
import subprocess
import sys

import pytest

from rfslam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(capi.lib().rfslam_b200_device_count() < 2, reason="needs two GPUs")
def test_two_gpu_allreduce_and_gather_match_one_gpu():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tools", "comm_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.count(": ok") == 2, out.stdout[-2000:] + out.stderr[-4000:]
