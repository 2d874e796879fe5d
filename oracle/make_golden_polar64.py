#!/usr/bin/env python
"""tests/golden/polar64_reference.npz: BASELINE.json configs[4] -- the 64-angle polar (alpha = -10 ... 21.5 deg, step 0.5, rans.h:54) of the
examples/conf.ini case (naca0012q_coarse -> naca0012q_mid FMG, implicit, tolerance 1e-4, <= 300 iterations per level) through the UNMODIFIED
reference's run_airfoil loop (oracle/_ref), cut into 8 warm-started chains of 8 angles exactly as `bench.py --workload polar64 --gpus 8` shards it.

    python oracle/make_golden_polar64.py chain K OUT.npz     one chain (run the eight in parallel, ~1 core each)
    python oracle/make_golden_polar64.py merge DIR           -> tests/golden/polar64_reference.npz

TEST INFRASTRUCTURE ONLY (needs /root/reference or the prebuilt oracle/_ref)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALPHAS = [-10.0 + 0.5 * k for k in range(64)]
MESHES = ["/root/reference/examples/rans/naca0012q_coarse.msh", "/root/reference/examples/rans/naca0012q_mid.msh"]

if sys.argv[1] == "chain":
    from oracle import ref
    k = int(sys.argv[2])
    mine = ALPHAS[8 * k:8 * k + 8]
    t0 = time.time()
    r = ref.run_sweep(MESHES, mine, implicit=True, relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, tolerance=1e-4,
                      rhs_iterations=5, max_iterations=300)
    np.savez(sys.argv[3], alphas=np.array(mine), cl=r["cl"], cd=r["cd"], cm=r["cm"], iters=r["iters"], seconds=time.time() - t0)
    print(k, mine, r["iters"], "%.1f s" % (time.time() - t0))
else:
    # chains that have finished (the two chains at alpha >= 14 deg need hours on one core each: stalled inviscid flow, the
    # reference's outer iteration runs to its 300-iteration limit on both levels with 6 x <= 500 GMRES iterations in each)
    parts = [np.load(os.path.join(sys.argv[2], "chain%d.npz" % k)) for k in range(8) if os.path.exists(os.path.join(sys.argv[2], "chain%d.npz" % k))]
    out = {n: np.concatenate([p[n] for p in parts]) for n in ("alphas", "cl", "cd", "cm", "iters")}
    out["chain_seconds"] = np.array([float(p["seconds"]) for p in parts])
    np.savez(os.path.join(ROOT, "tests", "golden", "polar64_reference.npz"), **out)
    for a, cl, cd, cm, it in zip(out["alphas"], out["cl"], out["cd"], out["cm"], out["iters"]):
        print("%6.1f  % .8f  % .8f  % .8f  %d" % (a, cl, cd, cm, it))
