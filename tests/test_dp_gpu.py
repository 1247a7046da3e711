"""Two-rank data-parallel parity: tools/dp_parity.py under torchrun.  With >= 2 CUDA devices the ranks run one per GPU
over NCCL; on a single-GPU box both ranks share cuda:0 and the collectives run over gloo (same host logic, same kernels),
so the N>1 path is exercised wherever the GPU tests run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("arch,unbalanced,pipeline", [("proton", False, 0), ("neutron", False, 0), ("proton", True, 0),
                                                     ("neutron", True, 0), ("proton", True, 1), ("neutron", False, 1)])
def test_two_rank_step_equals_global_batch_step(arch, unbalanced, pipeline):
    """neutron additionally exercises SyncBN (BatchNorm partial sums all-reduced between the reduce and apply kernels);
    ``unbalanced``: rank 1 holds no row of expert 0 although the expert is alive globally — Adam, spectral-norm u/v and
    BatchNorm running statistics must still advance there (replicas bit-identical after the step)."""
    port = 29611 + 2 * ["proton", "neutron"].index(arch) + 4 * int(unbalanced) + 8 * pipeline
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_parity.py"), arch]
    if unbalanced:
        cmd.append("--unbalanced")
    if torch.cuda.device_count() < 2:
        cmd.append("--one-gpu")
    # pipeline: the generator's Adam runs rectangle by rectangle behind the chunked all-reduce of fc2's bucket
    # (MoEWrapper._adam_pipelined); the replicas must stay bit-identical and the gradients at the same bounds
    env = dict(os.environ, ES_DP_PIPELINE_ADAM=str(pipeline))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "replicas bit-identical: True" in r.stdout
    assert f"pipelined Adam rectangles: {3 if pipeline else 0}" in r.stdout
    if unbalanced:
        assert "a rank holds no row of live expert 0: True" in r.stdout
