"""Multi-process tests: world_size 2 over gloo on CPU (host plumbing), and 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(mode, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), mode]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)


def test_two_ranks_gloo_host_plumbing():
    r = _torchrun("cpu", 2, 29631)
    assert "MP_CPU_OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_two_gpus_slab_run_matches_reference():
    from fastpm_b200 import _lib
    lib = _lib.load()
    if lib.fpm_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun("gpu", 2, 29641)
    assert "MP_GPU_OK" in r.stdout, r.stdout[-4000:]
