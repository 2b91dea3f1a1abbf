"""Regenerates the committed fixtures under tests/golden/ from the compiled reference (oracle/_ref).

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
  lightcone_deltak.npz   delta_k of tests/lightcone.lua (nc=64, box 512, seed 100, remove_cosmic_variance,
                         powerspec.txt) produced by the reference's fastpm_ic_fill_gaussiank + induce_correlation,
                         float32 in the reference's untransposed layout [x][y][N/2+1] complex.
  small_run.npz          a complete small reference run (nc=16, B=2, 4 steps, fastpm mode): IC delta_k, particle
                         state after 2LPT and after evolve, P(k) per force evaluation.
  powerspec.txt          the reference's tests/powerspec.txt (input linear P(k) table), copied verbatim.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402

pk = open(os.path.join(HERE, "powerspec.txt")).read()

s = ref.Session(nc=64, boxsize=512, pm_nc_factor=1, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0, compute_potential=True)
dk, var, s8 = s.ic_deltak(100, pk, remove_variance=True)
np.savez_compressed(os.path.join(HERE, "lightcone_deltak.npz"), delta_k=dk)
s.close()

s = ref.Session(nc=16, boxsize=32.0, pm_nc_factor=2, force_mode="fastpm", growth_mode="LCDM", np_alloc_factor=2.0)
dk, _, _ = s.ic_deltak(7, pk)
s.setup_lpt(dk, 0.1)
p0 = s.get_particles()
steps = np.linspace(0.1, 1.0, 4)
s.evolve(steps)
p1 = s.get_particles()
recs = s.records()
np.savez_compressed(os.path.join(HERE, "small_run.npz"), delta_k=dk, steps=steps,
                    x0=p0["x"], v0=p0["v"], id=p0["id"], x1=p1["x"], v1=p1["v"],
                    pk_k=np.array([r["k"] for r in recs]), pk_p=np.array([r["p"] for r in recs]),
                    pk_n=np.array([r["nmodes"] for r in recs]), vel_std=np.array([r["vel_std"] for r in recs]))
s.close()
print("fixtures written")
