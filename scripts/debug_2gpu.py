"""Two-rank diagnostic of the CUDA-IPC arena set-up (run with torch.distributed.run, 2 processes)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from fastpm_b200 import _lib, multigpu
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
use_nccl = os.environ.get("DBG_NCCL", "1") == "1"
dist.init_process_group(backend="nccl" if use_nccl else "gloo", **({"device_id": torch.device("cuda", local)} if use_nccl else {}))
lib = _lib.require_device(local)
def state(tag, ptr=None):
    out = (C.c_int * 4)()
    lib.fpm_debug_state(C.c_void_p(ptr), out)
    print("rank %d %-28s current=%d lib=%d ptr_device=%d type=%d" % (rank, tag, out[0], out[1], out[2], out[3]), flush=True)
state("after require_device")
multigpu.init_comm(lib)
lib.fastpm_b200_arena_base.restype = C.c_void_p
base = lib.fastpm_b200_arena_base()
state("after comm_init (arena)", base)
rc = lib.fpm_memset(C.c_void_p(base + (1 << 20)), 0, 4096)
print("rank %d memset rc=%d %s" % (rank, rc, lib.fpm_last_error().decode() if rc else ""), flush=True)
rc = lib.fpm_sync()
print("rank %d sync rc=%d" % (rank, rc), flush=True)
rc = lib.fpm_xbarrier(); rc2 = lib.fpm_sync()
print("rank %d xbarrier rc=%d sync rc=%d %s" % (rank, rc, rc2, lib.fpm_last_error().decode() if (rc or rc2) else ""), flush=True)
dist.barrier()
