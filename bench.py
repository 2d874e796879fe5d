#!/usr/bin/env python
"""bench.py -- RANS cell-updates/s of the explicit pseudo-time iteration on B200.

Contract (driver): python bench.py --gpus N --steps K --warmup W  (torchrun for N>1), ONE JSON line on rank 0.

A "step" is one explicitSolver::solve() (local dt + 3 RK stages, each = wall-ghost copy, limiter, face flux,
gather + update; the gradients of the iteration-start state are computed once) over the whole mesh.
Workload: BASELINE.json configs[2], the synthetic 16M-cell mixed tri/quad NACA0012 O-mesh (the mesh the metric and the
">= 60 % of the HBM roofline" target are quoted on), "RANS-SA" (which in the reference is the Roe flux + no-slip wall +
always-on gradients), second order, Green-Gauss, limiter_k 5, relaxation 0.9, CFL 1.5, perturbed free
stream (default_rng(12345), 1e-3).  --gpus N cuts the SAME 16M mesh into N pieces (strong scaling, configs[2]);
--scaling weak runs 8M cells per GPU (configs[3]: 64M cells on 8 GPUs); --workload synthetic-1M-mixed-omesh is configs[1].
Every N > 1 run carries a parity probe: rank 0 also runs the first 10 iterations of the same mesh un-partitioned and the
residual-norm histories must agree to 1e-10 relative ("parity" in the JSON line).

 value      cell-updates/s, state resident in HBM, CUDA events on the solver's stream around the K steps
 e2e        the same metric through the C ABI with HOST state every step: afx_rans_set_q (pinned H2D) ->
            afx_rans_step_explicit -> afx_rans_get_q (D2H) + the norm
 roofline   the face-flux kernel k_flux: algorithmic bytes (144*N + 48*E per launch) / its mean
            launch time, measured with CUDA events between the phases right after the timed region
 cpu_baseline  the unmodified reference (oracle/_ref) timed on this host on a bounded sample of the same workload
 --impl reference  that CPU reference as its own arm
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name: (ni, nj, n_quad_layers) -> SURVEY.md 8d
    "synthetic-64k-mixed-omesh": (256, 160, 64),
    "synthetic-1M-mixed-omesh": (1024, 640, 256),
    "synthetic-2M-mixed-omesh": (2048, 640, 256),
    "synthetic-4M-mixed-omesh": (2048, 1280, 512),
    "synthetic-8M-mixed-omesh": (4096, 1280, 512),
    "synthetic-16M-mixed-omesh": (4096, 2560, 1024),
    "synthetic-32M-mixed-omesh": (8192, 2560, 1024),
    "synthetic-64M-mixed-omesh": (8192, 5120, 2048),
}
HEADLINE = "synthetic-16M-mixed-omesh"  # BASELINE.json configs[2]: 1/2/4/8 GPUs on this one mesh (strong scaling)
# --scaling weak, BASELINE.json configs[3]: 8M cells per GPU (64M cells on 8 GPUs)
WEAK = {1: "synthetic-8M-mixed-omesh", 2: "synthetic-16M-mixed-omesh", 4: "synthetic-32M-mixed-omesh", 8: "synthetic-64M-mixed-omesh"}
PARITY_ITERS, PARITY_RTOL = 10, 1e-10
BCS = {"farfield": ("farfield", dict(mach=0.2, angle=1.0 * 0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
VISC, GRAD, SECOND, LIMK, CFL, RELAX = "spallart-allmaras", "green-gauss", True, 5.0, 1.5, 0.9
CPU_SAMPLE = "synthetic-64k-mixed-omesh"


def perturbed(q, N):
    rng = np.random.default_rng(12345)
    q = q.copy()
    q[:4 * N] *= 1.0 + 1e-3 * rng.uniform(-1, 1, 4 * N)
    return q


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """N > 1: keep this rank -- and the pinned host buffers it allocates for the end-to-end leg -- on the CPUs next to its GPU
    (nvmlDeviceGetCpuAffinity).  torchrun does not place its ranks: with 8 of them moving 2 x 67 MB per step at once, buffers on
    the far socket go through the inter-socket links (round 1: e2e at N = 8 was 2.5x N = 1 for 1/8 of the bytes per rank).
    Returns the number of CPUs the rank may use, or None when the topology cannot be read (nothing is changed then).
    AFX_BENCH_NUMA=0 disables."""
    if os.environ.get("AFX_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",")]
            if not all(v.isdigit() for v in ids):
                return None
            idx = int(ids[local])
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx), ((os.cpu_count() or 64) + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
        if len(cpus) < max(4, len(allowed) // 4):  # a container that shows only a sliver of the GPU's socket: leave the rank where it is
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def reference_cpu(steps, warmup, sample=CPU_SAMPLE):
    """The reference's own explicitSolver on the host cores (oracle/_ref), bounded sample of the workload."""
    import aeroflex_b200 as afx
    from oracle import ref
    import tempfile
    ni, nj, nq = WORKLOADS[sample]
    m = afx.Mesh.synth_omesh(ni, nj, nq, 150.0)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "sample.msh")
        m.write_msh(path)
        rm = ref.RefMesh(path)  # rans::mesh reads it, as in the reference
    s = ref.RefSolver(rm, False, VISC)
    s.set_bcs(BCS); s.set_options(SECOND, GRAD, LIMK, CFL); s.init(); s.refill_bcs()
    s.set("q", perturbed(s.get("q"), rm.N))
    for _ in range(warmup):
        s.explicit_solve(RELAX)
    t0 = time.perf_counter()
    for _ in range(steps):
        s.explicit_solve(RELAX)
    dt = time.perf_counter() - t0
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": rm.N * steps / dt, "unit": "cell-updates/s", "cores": threads, "kind": "reference",
            "sample": "%s (N=%d, E=%d): %d explicit iterations (after %d warm-up) of the unmodified reference headers (Eigen-API stand-in, "
                      "-O3 -fopenmp, OMP_NUM_THREADS=%d; the reference's face loops are serial, so 1 core does the hot loops), %.1f s; "
                      "cell-updates/s is size-normalised: the serial reference does not get faster per cell on a larger mesh"
                      % (sample, rm.N, rm.E, steps, warmup, threads, dt)}, dt / steps * 1e3


def port_cpu(steps):
    """The restated oracle, threaded variant (-O3 -march=native -fopenmp), on all host cores: the strongest CPU number
    we can produce for this path (bit-identical to the serial restatement; the reference itself cannot thread its face loops)."""
    import aeroflex_b200 as afx
    from oracle import orc
    ni, nj, nq = WORKLOADS[CPU_SAMPLE]
    m = afx.Mesh.synth_omesh(ni, nj, nq, 150.0)
    x, y, cells, b0, b1 = m.elements()
    om = orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names, fast=True)
    o = orc.OracleSolver(om, viscosity=VISC, fast=True)
    o.set_bcs(BCS); o.set_options(SECOND, GRAD, LIMK, CFL); o.init(); o.refill_bcs()
    o.q[:] = perturbed(o.q.copy(), m.N)
    o.explicit_solve_omp(RELAX)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.explicit_solve_omp(RELAX)
    dt = time.perf_counter() - t0
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": m.N * steps / dt, "unit": "cell-updates/s", "cores": threads, "kind": "port",
            "sample": "%s: %d iterations of oracle/rans_oracle.c orc_explicit_solve_omp, %d threads, %.1f s" % (CPU_SAMPLE, steps, threads, dt)}


def shared_config(workload, world, weak):
    """The part of `config` BOTH arms print, so that the two lines describe one configuration by construction."""
    ni, nj, nq = WORKLOADS[workload]
    return {"workload": workload, "cells": ni * (nq + 2 * (nj - nq)), "mesh": "NACA0012 O-mesh %d x %d, inner %d layers quads, outer layers split into triangles" % (ni, nj, nq),
            "scheme": "explicit 3-stage RK, 2nd order MUSCL, Green-Gauss, Venkatakrishnan k=5",
            "viscosity": "spallart-allmaras (reference semantics: Roe flux + no-slip wall + gradients, SURVEY F1/F2)",
            "cfl": CFL, "relaxation": RELAX, "gpus": world, "scaling": "weak (8M cells per GPU)" if weak else "strong (one mesh cut into N pieces)",
            "cpu_arms": "the CPU reference (--impl reference, cpu_baseline) is timed on a bounded sample of the same mesh family (%s) -- its mesh "
                        "reader and serial face loops need minutes per iteration at this size; cell-updates/s is size-normalised" % CPU_SAMPLE,
            "l2": "working set far larger than the 126 MB L2 (inputs larger than L2, no flush needed)"}


# ---- BASELINE.json configs[0] and configs[4]: the airfoil polar of examples/conf.ini through Rans::run_airfoil ----
POLAR = {"confini-polar": [1.0, 4.0, 7.0],                             # conf.ini: alpha_start 1, alpha_end 7, alpha_step 3
         "polar64": [-10.0 + 0.5 * k for k in range(64)]}              # alpha -10 ... 21.5 in steps of 0.5 (rans.h:54 loop rule)
POLAR_SET = dict(implicit=True, relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, tolerance=1e-4, rhs_iterations=5, max_iterations=300)


def prolongation_csr(coarse, fine):
    """multigrid::gen_mapper (multigrid.h:100-178) for the benchmark harness: all pairs, numpy.  (The product builds these
    weights in its C++ adapter with a k-d tree, host/rans_b200/multigrid.h; tests/test_cpp_host.py pins both to the reference.)"""
    xc, yc, ac = np.array(coarse.ccx), np.array(coarse.ccy), np.array(coarse.area)
    xf, yf = np.array(fine.ccx), np.array(fine.ccy)
    row_begin = np.zeros(len(xf) + 1, np.uint32)
    cols, ws = [], []
    for i in range(len(xf)):
        d2 = (xf[i] - xc) * (xf[i] - xc) + (yf[i] - yc) * (yf[i] - yc)
        j = np.flatnonzero(d2 < 2 * ac)
        si = 1 / np.maximum(0.1 * np.sqrt(ac[j]), np.sqrt(d2[j]))
        scale = 0.0
        for v in si:
            scale += v
        cols.append(j.astype(np.uint32)); ws.append(si / scale)
        row_begin[i + 1] = row_begin[i] + len(j)
    return row_begin, np.concatenate(cols), np.concatenate(ws)


def polar_meshes(afx):
    """naca0012q_coarse + naca0012q_mid (rans.h:84-85) from the committed fixtures (the GPU box has no /root/reference)."""
    out = []
    for tag in ("naca0012q_coarse_euler_gg_o2", "naca0012q_mid_mesh"):
        d = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"), allow_pickle=False)
        out.append(afx.Mesh.from_elements(d["x"], d["y"], d["cells"], d["is_tri"], d["b0"], d["b1"], d["bpatch"], [str(n) for n in d["patch_names"]]))
    return out


def polar_main(a, rank, world, local):
    """metric: angles of attack per second, whole job; every rank runs a contiguous chain of the angles (replicas, no communication)."""
    alphas = POLAR[a.workload]
    config = {"workload": a.workload, "case": "examples/conf.ini: naca0012q_coarse -> naca0012q_mid FMG, implicit, inviscid, slip wall, M = 0.2, CFL 40 -> 100, "
              "relaxation 0.9, tolerance 1e-4, <= 300 iterations per level", "angles": len(alphas), "gpus": world,
              "sharding": "contiguous warm-started chains of angles, one chain per GPU, no communication (replicas only)"}
    if a.impl == "reference":
        if rank != 0:
            return
        import tempfile
        import aeroflex_b200 as afx
        from oracle import ref
        n_ref = max(1, min(a.steps, 2))  # bounded: ~70 s per angle
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for k, m in enumerate(polar_meshes(afx)):
                paths.append(os.path.join(td, "level%d.msh" % k)); m.write_msh(paths[-1])
            t0 = time.perf_counter()
            r = ref.run_sweep(paths, alphas[:n_ref], **{k: v for k, v in POLAR_SET.items()})
            dt = time.perf_counter() - t0
        cpu = {"value": n_ref / dt, "unit": "angles/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)), "kind": "reference",
               "sample": "the first %d angles of the polar through the unmodified reference's run_airfoil loop, %.1f s" % (n_ref, dt)}
        print(json.dumps({"impl": "reference", "metric": "airfoil polar angles/s", "value": cpu["value"], "unit": "angles/s", "n_gpus": a.gpus, "steps": n_ref,
                          "warmup": 0, "ms_per_step": dt / n_ref * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "shipped meshes (fixtures)", "config": config, "cpu_baseline": cpu,
                          "e2e": {"value": cpu["value"], "unit": "angles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "forces": {"alpha": alphas[:n_ref], "cl": list(r["cl"]), "cd": list(r["cd"]), "cm": list(r["cm"])}}))
        return
    import torch
    import aeroflex_b200 as afx
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    meshes = polar_meshes(afx)
    csr = prolongation_csr(meshes[0], meshes[1])
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}

    def make():
        lv = [afx.GpuSolver(m, device=local, math=a.math) for m in meshes]
        for s in lv:
            s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0)
        return lv, afx.Prolongation(lv[0], lv[1], *csr)
    lo, hi = rank * len(alphas) // world, (rank + 1) * len(alphas) // world
    mine = alphas[lo:hi]
    lv, P = make()
    afx.sweep_fmg(lv, [P], mine[:1], **POLAR_SET)  # warm-up: one angle (allocations, graph-free implicit path)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = sum(s.launch_count() for s in lv)
    t0 = time.perf_counter()
    r = afx.sweep_fmg(lv, [P], mine, **POLAR_SET)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = sum(s.launch_count() for s in lv) - l0
    rows = [dict(alpha=al, cl=float(r["cl"][k]), cd=float(r["cd"][k]), cm=float(r["cm"][k]), iterations=int(r["iterations"][k]), residual=float(r["residual"][k]))
            for k, al in enumerate(mine)]
    allrows, t_max = [rows], dt
    if dist is not None:
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    parity = None
    if rank == 0:  # converged forces of the angles the reference fixture holds (both sides at 1e-10), a fresh chain
        g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_naca0012q_fmg.npz"), allow_pickle=False)
        lv2, P2 = make()
        rp = afx.sweep_fmg(lv2, [P2], list(g["alphas"]), **dict(POLAR_SET, tolerance=1e-10, max_iterations=400))
        rel = {k: float(np.max(np.abs(rp[k] - g[k]) / np.abs(g[k]))) for k in ("cl", "cd", "cm")}
        parity = {"what": "CL/CD/CM at alpha = %s, both sides driven to 1e-10 (tests/golden/sweep_naca0012q_fmg.npz, the unmodified reference)" % list(g["alphas"]),
                  "max_rel_diff": rel, "ok": bool(rel["cl"] <= 1e-7 and rel["cd"] <= 1e-6 and rel["cm"] <= 1e-6),
                  "note": "the reference's own ILUT/GMRES iteration stagnates near 1e-11 on this case: 1e-7 / 1e-6 is what its fixture supports; the 1e-8 force "
                          "tolerance is asserted against the explicit-path fixture (tests/golden/converged_naca0012q_coarse_explicit.npz)"}
    if rank == 0:
        flat = [x for rws in allrows for x in rws]
        if parity is not None and len(alphas) == 64:
            # the whole polar against the unmodified reference's own run_airfoil loop on the same angles (tests/golden/polar64_reference.npz,
            # 8 warm-started chains of 8, both sides stopped at 1e-4 with different linear solvers): the attached-flow range only --
            # beyond 14 deg the inviscid flow stalls and both iterations run to their limits
            gp = np.load(os.path.join(ROOT, "tests", "golden", "polar64_reference.npz"), allow_pickle=False)
            ref_by_alpha = {float(x): k for k, x in enumerate(gp["alphas"])}
            sel = [(r_, ref_by_alpha[r_["alpha"]]) for r_ in flat if r_["alpha"] <= 13.5 and r_["alpha"] in ref_by_alpha]
            parity["polar_vs_reference"] = {
                "angles_compared": len(sel), "tolerance_both_sides": 1e-4,
                "max_abs_diff": {k: float(max(abs(r_[k] - float(gp[k][j])) for r_, j in sel)) for k in ("cl", "cd", "cm")},
                "iterations": {"ours": int(sum(r_["iterations"] for r_, _ in sel)), "reference": int(sum(int(gp["iters"][j]) for _, j in sel))},
                "same_chains": world == 8}
        print(json.dumps({"metric": "airfoil polar angles/s", "value": len(alphas) / t_max, "unit": "angles/s", "n_gpus": world, "steps": len(alphas), "warmup": 1,
                          "ms_per_step": t_max / max(1, len(mine)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "shipped meshes (fixtures)", "config": config, "gpu_launches": int(launches), "seconds": t_max,
                          "e2e": {"value": len(alphas) / t_max, "unit": "angles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 24,
                                  "path": "afx_rans_sweep_fmg: the states stay on the device, CL/CD/CM come back per angle"},
                          "parity": parity, "polar": flat}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="default: %s (BASELINE.json configs[2]); any key of WORKLOADS" % HEADLINE)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="strong: the workload cut into N pieces; weak: 8M cells per GPU (configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-probe", action="store_true", help="N > 1: skip the un-partitioned 10-iteration run on rank 0")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--fuse-lim0", type=int, default=1, help="1 (default): k_dt_grad also writes the first stage's limiters; 0: separate k_limiter launch")
    ap.add_argument("--fused", type=int, default=0, help="1: one fused kernel per RK stage on shared-memory tiles; 0 (default): limiter / flux / gather kernels")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    weak = a.scaling == "weak" and a.workload is None
    if a.workload in POLAR:
        return polar_main(a, rank, world, local)
    workload = a.workload or (WEAK.get(world) if weak else HEADLINE)
    if workload not in WORKLOADS:
        raise SystemExit("bench.py: unknown workload %r" % workload)
    scaling = "weak" if weak else "strong"
    config = shared_config(workload, world, weak)

    if a.impl == "reference":
        if rank != 0:
            return
        # K steps after W warm-up steps exactly as asked; every step is one explicit iteration of the unmodified reference on
        # the bounded sample (~0.2 s per iteration on 64k cells)
        steps = max(1, a.steps)
        cpu, ms = reference_cpu(steps, a.warmup)
        print(json.dumps({"impl": "reference", "metric": "RANS cell-updates/s", "value": cpu["value"], "unit": "cell-updates/s",
                          "n_gpus": a.gpus, "steps": steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                                                       "d2h_bytes_per_step": 0}}))
        return

    if a.fused:
        os.environ["AFX_FUSED"] = "1"  # the tiles (and the graph-bisection numbering) are built at creation
    os.environ["AFX_FUSE_LIM0"] = "1" if a.fuse_lim0 else "0"
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None  # before the library starts its OpenMP threads
    import torch
    import aeroflex_b200 as afx
    if afx.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device visible; the product has no CPU path")
    if afx.is_emulation():
        raise SystemExit("bench.py: AFX_LIB names the host emulation of tests/emu; only the CUDA library is measured")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_setup = time.perf_counter()
    ni, nj, nq = WORKLOADS[workload]
    mesh = afx.Mesh.synth_omesh(ni, nj, nq, 150.0)
    N, G, E = mesh.N, mesh.G, mesh.E
    part = None
    if world > 1:
        part = afx.Partition(mesh, world, rank)
        ids = [afx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        s = afx.GpuSolver(part, viscosity=VISC, device=local, math=a.math, nccl_id=ids[0])
        if os.environ.get("AFX_HALO", "p2p") == "p2p":  # halo through NVLink peer memory, fused into the update kernel
            blobs = [None] * world
            dist.all_gather_object(blobs, s.p2p_export())
            s.p2p_connect(blobs)
    else:
        s = afx.GpuSolver(mesh, viscosity=VISC, device=local, math=a.math)
    s.set_bcs(BCS); s.set_options(SECOND, GRAD, LIMK, CFL); s.init(); s.refill_bcs()
    s.set_fused(a.fused)
    tiles = s.tile_info()
    details = {"math": a.math + (" (shared reciprocals + FMA; parity 1e-10 tested)" if a.math == "fast" else " (bit-identical to the CPU reference)"),
               "edges": E, "ghost_cells": G, "setup_s": None}
    details["stage_kernel"] = ("fused k_stage: %d tiles of <= %d cells, %d B shared memory, %d CTAs/SM, staging overhead %.2fx"
                               % (tiles["tiles"], tiles["tile_cells"], tiles["smem_bytes"], tiles["ctas_per_sm"],
                                  tiles["local_cells"] / max(1, (N if part is None else part.n_own)))) if tiles["fused"] else ("k_dt_grad writes the first stage's limiters; k_limiter (stages 2, 3) + k_flux + k_gather_update" if (a.fuse_lim0 and SECOND) else "k_limiter + k_flux + k_gather_update")
    base = np.zeros(4 * (N + G))
    s.get_q(base)  # a partitioned solver fills its own entries of the global vector
    q0 = perturbed(base, N)  # multiplicative, by GLOBAL index: every rank perturbs its own entries exactly as one GPU would
    del base
    if world > 1:
        details["numa"] = ("rank 0 bound to the %d CPUs next to its GPU (pinned e2e buffers allocated there)" % numa_cpus) if numa_cpus \
            else "ranks not bound (topology unavailable or AFX_BENCH_NUMA=0)"
    details["parallelism"] = "1 GPU" if world == 1 else \
        "domain decomposition: %d %s partitions, 2-layer halo refreshed every RK stage, halo=%s (this rank: %d owned + %d halo cells, %d peers)" \
        % (world, os.environ.get("AFX_PARTITION", "hilbert"), s.halo_mode(), part.n_own, part.n_r1 + part.n_r2, part.n_peers)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # parity probe, part 1 (N > 1): the first PARITY_ITERS iterations of the PARTITIONED run from q0
    parity = None
    probe = world > 1 and not a.no_parity_probe
    if probe:
        s.set_q(q0)
        probe_norms = s.run(PARITY_ITERS, RELAX)
        probe_forces = np.array(s.wall_forces("wall"))
    s.set_q(q0)
    details["setup_s"] = time.perf_counter() - t_setup

    s.run(W, RELAX)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    l0 = s.launch_count()
    t0 = time.perf_counter()
    norms = s.run(a.steps, RELAX)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = s.last_device_ms()
    launches = s.launch_count() - l0
    clocks = sampler.finish()
    if not np.all(np.isfinite(norms)):
        raise SystemExit("bench.py: residual norm is not finite")
    t_ms = dev_ms
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    value = N * a.steps / (t_ms * 1e-3)  # N is the whole (partitioned) mesh

    # per-phase kernel times (CUDA events between the phases, same state, right after the timed region)
    prof = s.profile_explicit(5, RELAX)
    if dist is not None:  # how even the pieces are: min / max of every phase over the ranks, faces per rank
        allp = [None] * world
        dist.all_gather_object(allp, dict(prof, cells=part.n_own, faces=part.E, **{"halo_" + k: v for k, v in s.profile_halo_ms().items()}))
        details["rank_spread"] = {k: [min(p[k] for p in allp), max(p[k] for p in allp)] for k in allp[0]}
    n_loc = N if part is None else part.n_own
    share = n_loc / N  # this rank's share of the cells
    # ALGORITHMIC bytes per launch (they sum to 1488 N + 296 E per iteration however the phases are fused):
    #   dt 48N+32E, Green-Gauss 120N+48E, limiter 152N+24E, residual loop 184N+48E, stage update 104N.
    # The face kernel carries the residual loop's reads (144N+48E); the gather/update kernel the residual's qW + area (40N) and
    # the update (104N).  The flux buffer between them is an implementation temporary: it shows up in `traffic`, not here.
    if tiles["fused"]:
        kernels = {"k_stage": ((440.0 * N + 72.0 * E) * share, 3, prof["stage"]),
                   "k_dt_grad": ((168.0 * N + 80.0 * E) * share, 1, prof["dt_grad"])}
    else:
        lim0 = 1 if (a.fuse_lim0 and SECOND) else 0  # the first stage's limiter phase runs inside k_dt_grad
        kernels = {"k_flux": ((144.0 * N + 48.0 * E) * share, 3, prof["flux"]),
                   "k_limiter": ((152.0 * N + 24.0 * E) * share, 3 - lim0, prof["limiter"]),
                   "k_gather_update": (144.0 * N * share, 3, prof["gather_update"]),
                   "k_dt_grad": ((168.0 * N + 80.0 * E + lim0 * (152.0 * N + 24.0 * E)) * share, 1, prof["dt_grad"])}
    alg_iter = (1488.0 * N + 296.0 * E) * share
    assert abs(sum(b * n for b, n, _ in kernels.values()) - alg_iter) < 1e-6 * alg_iter
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(workload, {})
    except Exception:
        pass
    # What a kernel's OWN algorithm has to move when that is less than the SURVEY phase it stands for: in fast mode the limiters of
    # stages 2 and 3 read the stored projected extremes (64 B) instead of gx, gy, the centre offsets and the per-face geometry
    # (DESIGN.md 4/5): q 32 + extremes 64 + neighbour table 16 + lim 32 = 144 N.  The per-kernel fraction uses the smaller of the
    # two counts, so no kernel is credited with bytes it does not need; the ITERATION figure stays the SURVEY's 1488 N + 296 E.
    own = {}
    if not tiles["fused"] and a.math == "fast" and SECOND and os.environ.get("AFX_LIM_PM", "1") != "0":
        own["k_limiter"] = 144.0 * N * share
    per_kernel = {}
    for k, (bytes_, n_launch, ms) in kernels.items():
        t = ms / n_launch
        credited = min(bytes_, own.get(k, bytes_))
        per_kernel[k] = {"algorithmic_bytes_per_launch": credited, "survey_bytes_per_launch": bytes_, "launches_per_iteration": n_launch,
                         "kernel_ms": t, "achieved": credited / (t * 1e-3) / 1e9 if t > 0 else None,
                         "frac": credited / (t * 1e-3) / 1e9 / peak if t > 0 else None,
                         "share_of_iteration": ms / sum(v[2] for v in kernels.values()), "traffic": traffic_tab.get(k)}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["share_of_iteration"])
    roof = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["achieved"], "peak": peak, "unit": "GB/s",
            "frac": per_kernel[dom]["frac"], "traffic": per_kernel[dom]["traffic"],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "algorithmic_bytes_per_launch": per_kernel[dom]["algorithmic_bytes_per_launch"], "kernel_ms": per_kernel[dom]["kernel_ms"],
            "kernels": per_kernel, "phase_ms_per_iteration": prof,
            "iteration": {"algorithmic_bytes": alg_iter, "achieved": alg_iter / (t_ms / a.steps * 1e-3) / 1e9,
                          "frac": alg_iter / (t_ms / a.steps * 1e-3) / 1e9 / peak,
                          "frac_of_nominal_8TBs": alg_iter / (t_ms / a.steps * 1e-3) / 1e9 / 8000.0}}
    if not tiles["fused"] and prof["flux"] > 0:
        # BASELINE's "residual-loop HBM GB/s": the SURVEY's residual loop (184 N + 48 E) and stage update (104 N) are carried by
        # k_flux + k_gather_update together; one launch of each
        b = (288.0 * N + 48.0 * E) * share
        t = (prof["flux"] + prof["gather_update"]) / 3.0
        roof["residual_loop_and_update"] = {"algorithmic_bytes": b, "ms": t, "achieved": b / (t * 1e-3) / 1e9,
                                            "frac": b / (t * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": b / (t * 1e-3) / 1e9 / 8000.0}

    # end to end through the C ABI with host-resident state, pinned buffers
    # (a partitioned rank moves ITS piece: owned + halo + ghost rows, in its own numbering)
    nq4 = s.local_size4()
    hq = afx.pinned_array(nq4)
    s.get_q_local(hq)
    for _ in range(2):
        s.set_q_local(hq); s.solve(RELAX); s.get_q_local(hq)
    barrier()
    te = time.perf_counter()
    for _ in range(a.e2e_steps):
        s.set_q_local(hq)
        nr = s.solve(RELAX)
        s.get_q_local(hq)
    barrier()
    e2e_s = time.perf_counter() - te
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": N * a.e2e_steps / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": 8 * nq4 * world,
           "d2h_bytes_per_step": 8 * nq4 * world + 8, "steps": a.e2e_steps, "ms_per_step": e2e_s / a.e2e_steps * 1e3,
           "path": "afx_rans_set_q[_local](pinned host) -> afx_rans_step_explicit -> afx_rans_get_q[_local](pinned host) + norm"}

    # parity probe, part 2: rank 0 runs the same iterations on the un-partitioned mesh (its GPU holds both solvers)
    if probe:
        del hq
        if rank == 0:
            t1 = time.perf_counter()
            one = afx.GpuSolver(mesh, viscosity=VISC, device=local, math=a.math)
            one.set_bcs(BCS); one.set_options(SECOND, GRAD, LIMK, CFL); one.init(); one.refill_bcs()
            one.set_q(perturbed(one.get_q(), N))  # init() is bit-identical on every piece (tested): the same q0, now whole
            ref_norms = one.run(PARITY_ITERS, RELAX)
            ref_forces = np.array(one.wall_forces("wall"))
            del one
            rel = float(np.max(np.abs(probe_norms - ref_norms) / np.abs(ref_norms)))
            frel = float(np.max(np.abs(probe_forces - ref_forces) / np.maximum(np.abs(ref_forces), 1e-300)))
            parity = {"what": "first %d residual norms of the %d-GPU run vs the same mesh un-partitioned on rank 0's GPU (same math mode), "
                              "and CL/CD/CM after those iterations" % (PARITY_ITERS, world),
                      "norm_rtol": PARITY_RTOL, "norm_max_rel_diff": rel, "forces_max_rel_diff": frel,
                      "ok": bool(rel <= PARITY_RTOL and frel <= 1e-8), "probe_s": time.perf_counter() - t1}
        barrier()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu, _ = reference_cpu(20, 2)
        cpu_port = port_cpu(60)
    if rank == 0:
        out = {"metric": "RANS cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": a.steps, "warmup": W,
               "ms_per_step": t_ms / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "details": details, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roof, "wall_ms_per_step": wall_ms / a.steps, "final_residual_norm": float(norms[-1])}
        if parity is not None:
            out["parity"] = parity
        if cpu is not None:
            out["cpu_baseline"] = cpu
            out["cpu_port"] = cpu_port
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
