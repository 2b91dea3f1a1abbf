"""TEST INFRASTRUCTURE ONLY: runs bench.py's multi-GPU arm (fastpm_b200/multigpu.py: bench_main) under torchrun on the CPU against
the emulated library.  The arm talks to torch.cuda and NCCL for its plumbing; here those few calls are replaced (gloo process group,
CPU tensors, no-op device selection) so that everything else -- the Solver run on every rank, the timing logic, the JSON line -- is
the code the driver will launch on the GPUs.  Usage (tests/test_cpu_full_emulation.py):
    python -m torch.distributed.run --nproc-per-node N ... run_bench_multirank.py --gpus N --nc 8 --steps 3 --warmup 1"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)

import torch                                  # noqa: E402
import torch.distributed as dist              # noqa: E402
from fastpm_b200 import _lib                  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tests", "emul", "_build", "libfastpm_b200_emul.so")
os.environ.setdefault("FASTPM_B200_ARENA_GB", "0.25")
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.current_device = lambda: 0
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
dist.init_process_group(backend="gloo")
sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
