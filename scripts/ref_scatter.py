"""Run-to-run scatter of the REFERENCE itself at BASELINE configs[1] (nc = 256, 512^3 mesh, COLA, 10 steps): the compiled reference
(oracle/_ref) with two different OpenMP thread counts.  Its float32 deposit uses `omp atomic` (painter-cic.c:24), so the order
of additions -- and after ten steps of a chaotic system the positions -- depend on the thread count.

    python scripts/ref_scatter.py 8 && python scripts/ref_scatter.py 5 && python scripts/ref_scatter.py compare 8 5

Measured in this container (8 host cores, round 2): max 1.16e-4 Mpc/h, 99.99 % quantile 9.5e-6, 99.9999 % quantile 3.0e-5,
mean 5.6e-7 Mpc/h; P(k) 7e-8 relative; velocities 1.1e-5 of their maximum.  tests/test_gpu_c1.py sets its position tolerances
from these numbers."""
import sys, os, numpy as np
if sys.argv[1] == "compare":
    a, b = (np.load("/tmp/ref_scatter_%s.npz" % t) for t in sys.argv[2:4])
    L = 256.0
    d = np.abs(a["x"] - b["x"]); d = np.minimum(d, L - d).max(axis=1)
    print("max %.3g  q99.99 %.3g  q99.9999 %.3g  mean %.3g Mpc/h; P(k) %.3g; v %.3g" % (
        d.max(), np.quantile(d, 0.9999), np.quantile(d, 0.999999), d.mean(), np.abs(a["p"][:, 1:] / b["p"][:, 1:] - 1).max(),
        np.abs(a["v"] - b["v"]).max() / np.abs(a["v"]).max()))
    sys.exit(0)
os.environ["OMP_NUM_THREADS"]=sys.argv[1]
sys.path.insert(0,'/root/repo')
from oracle import ref
pk_text=open("/root/repo/tests/golden/powerspec.txt").read()
nc=256
s = ref.Session(nc=nc, boxsize=256.0, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0)
dk,_,_ = s.ic_deltak(100, pk_text)
s.setup_lpt(dk, 0.1)
s.evolve(np.linspace(0.1,1.0,10))
p=s.get_particles()
recs=s.records()
np.savez("/tmp/ref_scatter_%s.npz"%sys.argv[1], x=p["x"], v=p["v"], id=p["id"], p=np.array([r["p"] for r in recs]))
s.close()
print("done", sys.argv[1])
