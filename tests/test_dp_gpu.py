"""N-GPU data-parallel parity (needs >= 2 CUDA devices; skipped on a single-GPU box): tools/dp_parity.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("arch", ["proton", "neutron"])
def test_two_rank_step_equals_global_batch_step(arch):
    """neutron additionally exercises SyncBN: BatchNorm partial sums all-reduced between the reduce and apply kernels"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611" if arch == "proton" else "29613", os.path.join(ROOT, "tools", "dp_parity.py"), arch],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
