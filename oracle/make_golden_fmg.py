import os, sys, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref
R='/root/reference/examples/rans/'
t=time.time()
r = ref.run_sweep([R+"naca0012q_coarse.msh", R+"naca0012q_mid.msh"], [1.0, 4.0], implicit=True, tolerance=1e-10, max_iterations=400)
print(r, time.time()-t)
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'sweep_naca0012q_fmg.npz'), alphas=np.array([1.0,4.0]), cl=r["cl"], cd=r["cd"], cm=r["cm"], iters=r["iters"], seconds=r["seconds"],
  meta=np.array("reference FMG sweep (rans.h:78-106): naca0012q_coarse -> naca0012q_mid, implicit, inviscid, green-gauss, second order, slip-wall, M=0.2, CFL 40->100, relaxation 0.9, tolerance 1e-10, 8 OpenMP threads"))
