"""-m gpu, needs >= 2 GPUs: launches tests/multi_gpu_check.py under torch.distributed.run (one
process per GPU, NCCL halo sums / all-reduces).  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("nproc,partition", [(2, "slabs"), (4, "slabs"), (4, "blocks"), (8, "blocks")])
def test_multi_gpu_parity(nproc, partition):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc),
           os.path.join(HERE, "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, SVFSI_PARTITION=partition))
    print(out.stdout[-4000:]); print(out.stderr[-2000:])
    assert out.returncode == 0 and "MULTI-GPU CHECK OK" in out.stdout
