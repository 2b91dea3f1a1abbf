"""One force evaluation + kick + drift at a given size, for ncu captures (not a benchmark)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fastpm_b200.solver import Solver, ForceEvent

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tab = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "powerspec.txt"))
g = Solver(nc=nc, boxsize=float(nc), pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=1.0)
g.setup_synthetic_ic(100, tab[:, 0], tab[:, 1], 0.1)


def on_force_after(solver_ptr, event_ptr, userdata):
    ev = C.cast(event_ptr, C.POINTER(ForceEvent)).contents
    g.powerspectrum_of(ev.pm, ev.delta_k)
    return 0


g.add_handler("FORCE", 1, on_force_after)
g.evolve(np.linspace(0.1, 1.0, 10)[:nsteps])
g.close()
