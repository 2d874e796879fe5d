#!/usr/bin/env python
"""BASELINE config 5: the 64-angle polar sweep (alpha = -10 ... 21.5 step 0.5, rans.h:54) of the conf.ini airfoil
case (implicit, FMG naca0012q coarse -> mid) that feeds the VLM viscous-correction database, sharded over the visible
GPUs as contiguous warm-start chains (scripts/polar_sweep.sh; replicas only, no communication).
   python scripts/polar_sweep_bench.py [n_gpus] [alpha_end]
Writes the meshes from the committed fixtures, runs the sweep and prints one JSON line with the wall time and the polar."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aeroflex_b200 as afx  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_cpp_host import CONF  # noqa: E402


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, afx.device_count())
    alpha_end = sys.argv[2] if len(sys.argv) > 2 else "21.5"
    with tempfile.TemporaryDirectory() as td:
        H.product_mesh(afx, H.load("naca0012q_coarse_euler_gg_o2")).write_msh(os.path.join(td, "naca0012q_coarse.msh"))
        H.product_mesh(afx, H.load("naca0012q_mid_mesh")).write_msh(os.path.join(td, "naca0012q_mid.msh"))
        conf = CONF % dict(solver="implicit", tol="1e-4", max_it=300, alpha_end=alpha_end, start_cfl="40.0")
        conf = conf.replace("alpha_start = 1.0", "alpha_start = -10.0").replace("alpha_step = 3.0", "alpha_step = 0.5")
        ini = os.path.join(td, "conf.ini")
        open(ini, "w").write(conf)
        t0 = time.perf_counter()
        out = subprocess.run(["bash", os.path.join(ROOT, "scripts", "polar_sweep.sh"), ini, td + "/", str(n_gpus)], capture_output=True, text=True)
        wall = time.perf_counter() - t0
    rows = [[float(v) for v in l.split()] for l in out.stdout.splitlines() if l and not l.startswith("#")]
    print(json.dumps({"config": "polar sweep, conf.ini airfoil case (implicit, FMG naca0012q coarse->mid, tol 1e-4, <=300 outer iterations per level)",
                      "n_gpus": n_gpus, "alphas": len(rows), "wall_s": wall, "alphas_per_s": len(rows) / wall if wall > 0 else None,
                      "polar_alpha_cl_cd_cm": rows, "rc": out.returncode, "stderr_tail": out.stderr[-300:]}))


if __name__ == "__main__":
    main()
