"""Multi-GPU parity of the partitioned path: runs tests/dist_check.py under torchrun on 2 GPUs when the box has them
(the single-GPU round-end box skips it; the host-side partition logic is covered on CPU by test_partition_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mesh", ["12", "eggshell"])  # structured cube (slabs) / unstructured numbering (Cuthill-McKee ordering)
def test_partitioned_two_gpus_match_single_gpu(mesh):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "dist_check.py"), mesh]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0 and "DIST_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
