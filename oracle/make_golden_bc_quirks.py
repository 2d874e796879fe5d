"""Golden vectors for solver::get_boundary_variables (solver.h:597-611), made with the UNMODIFIED reference headers
(oracle/_ref): the search over the boundary edges stops at the first "farfield" edge -- or at the first edge whose type is
the literal "inlet-outlet", whose variables are the defaults (M 0.2, angle 0).  On naca0012q_coarse the patch named
"farfield" comes first in the boundary list, so giving IT the type under test and the far-field type to the patch named
"wall" makes the search meet the type under test first.  Run here (needs /root/reference); writes tests/golden/bc_quirks.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

rm = ref.RefMesh("/root/reference/examples/rans/naca0012q_coarse.msh")
out = {}
for tag, typ in (("inlet_outlet", "inlet-outlet"), ("unknown", "something-else")):
    bcs = {"farfield": (typ, None), "wall": ("farfield", dict(mach=0.3, angle=0.05, T=1.0, p=1.0))}
    r = ref.RefSolver(rm)
    r.set_bcs(bcs); r.set_options(True, "green-gauss", 5.0, 1.2); r.init(); r.refill_bcs()
    out[tag + "_q_init"] = r.get("q").copy()
    out[tag + "_uniform_residual"] = np.array(r.uniform_residual())
    out[tag + "_norms"] = np.array([r.explicit_solve(0.9) for _ in range(3)])
    out[tag + "_q"] = r.get("q").copy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bc_quirks.npz"), **out)
print({k: (v.shape, float(np.ravel(v)[0])) for k, v in out.items()})
