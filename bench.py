#!/usr/bin/env python
"""bench.py — MLUPS of the fused D2Q9 collide+stream step on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path, rank 0 only)
    python bench.py --workload mrtcg_rt|rk_droplet|sedimentation|poiseuille|kbc_shear|csf_rt|cylinder_bb   (the other BASELINE.json configs)

A "step" is one lattice-Boltzmann time step of the whole grid.  Default workload at every N:
configs[1] of BASELINE.json — flow past a cylinder, D2Q9 BGK, compressible equilibrium,
immersed-boundary cylinder (multi-direct forcing), anti-bounce-back inlet/outlet rows, specular
side columns (test/cylinder_test.cpp of the reference), 8192 x 8192 nodes PER GPU (weak scaling:
the global grid is (8192 N) x 8192, slab-decomposed along axis 0 like test/decompose_domain.cpp).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))

FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md

W9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
CX9 = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=np.float64)
CY9 = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], dtype=np.float64)

# BASELINE.json configs -> workload; bytes/node = SURVEY §8(d)'s minimal-traffic model (fp64, every population
# read once and written once; two-phase models add their scalar moment planes)
WORKLOADS = {
    # configs[1] — the headline: test/cylinder_test.cpp at 8192 x 8192 per GPU
    "cylinder": dict(X=8192, Y=8192, bytes=144.0, nlat=1, kernel="k_bgk_interior<PULL,COMP,IBM>", driver="test/cylinder_test.cpp",
                     what="cylinder flow, D2Q9 BGK + IBM cylinder, ABB inlet/outlet, specular walls", cpu_sample=1024),
    # configs[1] read literally ("flow past cylinder ... with bounce-back"): the same free-stream channel with a staircase
    # cylinder under link-wise half-way bounce-back instead of the reference driver's immersed-boundary body
    "cylinder_bb": dict(X=8192, Y=8192, bytes=144.0, nlat=1, kernel="k_bgk_interior<PULL,COMP,NONE>",
                        driver="test/cylinder_test.cpp (boundary rows/columns) + the obstacle-wall rule of test/rectangle_sedimentation_test.cpp:186-196",
                        what="flow past a staircase cylinder, D2Q9 BGK compressible, half-way bounce-back links, ABB inlet/outlet, specular side columns",
                        cpu_sample=1024),
    # configs[0] at the parameters.toml grid (the 21 x 21 reference case is a parity test)
    "poiseuille": dict(X=2700, Y=2100, bytes=144.0, nlat=1, kernel="k_bgk_interior<PULL,INCOMP,NONE>",
                       driver="test/horizontal_poiseuille_test.cpp",
                       what="horizontal Poiseuille, D2Q9 BGK incompressible, pressure-periodic rows, bounce-back walls",
                       cpu_sample=1024),
    # configs[2]
    "mrtcg_rt": dict(X=16384, Y=16384, bytes=352.0, nlat=2, kernel="k_tp_fused<MRTCG,PIPE>",
                     driver="test/mrtcg_rayleigh_taylor.cpp",
                     what="MRT colour-gradient Rayleigh-Taylor (mrtcg-rayleigh-taylor-gamma3.toml), two lattices",
                     cpu_sample=512),
    # configs[3]
    "rk_droplet": dict(X=4096, Y=4096, bytes=304.0, nlat=2, kernel="k_tp_fused<RK,PIPE>",
                       driver="test/rk_static_droplet_test.cpp",
                       what="Rothman-Keller static droplet, R = L/4, two lattices", cpu_sample=512),
    # configs[4]
    "sedimentation": dict(X=4096, Y=8192, bytes=288.0, nlat=2, kernel="k_bgk_interior<PULL,COMP,NONE,ADE>",
                          driver="test/rectangle_sedimentation_test.cpp",
                          what="rectangle sedimentation: fluid + advection-diffusion lattice, bounce-back rectangle walls",
                          cpu_sample=1024),
    # configs[4] as BASELINE.json words it ("with immersed-boundary coupling (ibm)"): driver 15 has no ibm object (SURVEY §8
    # notes), so this is driver 15's loop plus a cylinder coupled the way driver 11 couples one; not a reference driver
    "sedimentation_ibm": dict(X=4096, Y=8192, bytes=288.0, nlat=2, kernel="k_bgk_interior<PULL,COMP,IBM,ADE>",
                              driver="test/rectangle_sedimentation_test.cpp + ibm as in test/cylinder_test.cpp",
                              what="rectangle sedimentation with an immersed cylinder: fluid (Guo-forced) + advection-diffusion lattice",
                              cpu_sample=1024),
    # not a BASELINE.json config: SURVEY §8(f) rank 2, the continuum-surface-force variant (three passes per step; bytes =
    # both colours' populations read + written once (288) + the carried interfacial tension read + written (32))
    "csf_rt": dict(X=8192, Y=8192, bytes=320.0, nlat=2, kernel="k_csf_collide_ring<PULL>",
                   driver="test/mrt_rayleigh_taylor.cpp",
                   what="MRT colour-gradient Rayleigh-Taylor with continuum surface force (curvature from nested 5x5 differences)",
                   cpu_sample=512),
    # not a BASELINE.json config: the SURVEY §8(f) rank-3 collision (ulbm::d2q9::kbc), fully periodic double shear layer
    "kbc_shear": dict(X=8192, Y=8192, bytes=144.0, nlat=1, kernel="k_bgk_interior<PULL,KBC>",
                      driver="test/ulbm_double_shear_flow.cpp",
                      what="double shear layer, D2Q9 entropic central-moment (KBC) collision, fully periodic", cpu_sample=512),
}

RED = dict(rho_0=3.0, alpha=0.7, A=0.5, nu=0.04, beta=0.7)      # configs/mrtcg-rayleigh-taylor-gamma3.toml
BLUE = dict(rho_0=1.0, alpha=0.1, A=0.5, nu=0.04, beta=-0.7)
RK_RED = dict(rho_0=1.2, alpha=1.0 / 3.0, A=1e-4, nu=0.16, beta=0.7)   # rk_static_droplet_test.cpp:504-506
RK_BLUE = dict(rho_0=1.0, alpha=0.2, A=1e-4, nu=0.14, beta=-0.7)
RT_FG = (6.25e-6, 0.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cylinder", choices=sorted(WORKLOADS),
                    help="cylinder = BASELINE.json configs[1] (the headline); the others are the remaining configs")
    ap.add_argument("--X", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--Y", type=int, default=0, help="columns (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="edge of the square grid the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.X = args.X or wl["X"]
    args.Y = args.Y or wl["Y"]
    args.cpu_sample = args.cpu_sample or wl["cpu_sample"]
    return args


# ---------------------------------------------------------------------------------------------
# workload description (shared by both arms)
# ---------------------------------------------------------------------------------------------
def lattice_parameters():
    """omega and u_lb exactly as params::lattice derives them from configs/parameters.toml"""
    import lbm_b200 as L

    p = L.params_from_toml(os.path.join(ROOT, "configs", "parameters.toml"), False)
    return p.omega, p.u


def channel_constants(H, W, u_max=0.1030985714):
    """test/horizontal_poiseuille_test.cpp:50-66"""
    tau = np.sqrt(3.0 / 16.0) + 0.5
    nu = (2.0 * tau - 1.0) / 6.0
    p_grad = 8.0 * nu * u_max / (W * W)
    return 1.0 / tau, 3.0 * (H - 1) * p_grad + 1.0, 1.0


def cylinder_markers(X, Y):
    """Cylinder of diameter ~X/9 (the reference's X = 9 l proportion, src/params.cpp:64) centred at
    (X/4, Y/2), markers about one lattice unit apart on the circle."""
    D = max(8.0, X / 9.0)
    n = max(16, int(round(np.pi * D)))
    th = 2.0 * np.pi * np.arange(n) / n
    return X / 4.0 + 0.5 * D * np.cos(th) + 0.37, Y / 2.0 + 0.5 * D * np.sin(th) + 0.21


def cylinder_solid(Xg, Y, X1):
    """GLOBAL {Xg,Y} solid mask of the staircase cylinder: same centre and diameter as cylinder_markers of the first slab"""
    D = max(8.0, X1 / 9.0)
    solid = np.zeros((Xg, Y), dtype=np.uint8)
    lo, hi = max(0, int(X1 / 4.0 - D)), min(Xg, int(X1 / 4.0 + D) + 1)
    r2 = (np.arange(lo, hi)[:, None] - (X1 / 4.0 + 0.37)) ** 2 + (np.arange(Y)[None, :] - (Y / 2.0 + 0.21)) ** 2
    solid[lo:hi] = r2 <= (0.5 * D) ** 2
    return solid


def sedimentation_body(X, Y):
    """sedimentation_ibm: a cylinder of diameter ~X/9 upstream of the rectangle (the flow runs along axis 1), centred at
    (X/2, Y/4) of the first slab, markers about one lattice unit apart"""
    D = max(8.0, X / 9.0)
    n = max(16, int(round(np.pi * D)))
    th = 2.0 * np.pi * np.arange(n) / n
    return X / 2.0 + 0.5 * D * np.cos(th) + 0.37, Y / 4.0 + 0.5 * D * np.sin(th) + 0.21


def write_markers_toml(path, xs, ys):
    with open(path, "w") as fh:
        fh.write("[cylinder-a]\n")
        fh.write("x = [" + ", ".join(repr(float(v)) for v in xs) + "]\n")
        fh.write("y = [" + ", ".join(repr(float(v)) for v in ys) + "]\n")


def sedimentation_geometry(X, Y):
    """rectangle_sedimentation_test.cpp:73-75,89-95 scaled from the parameters.toml grid (2700 x 2100)"""
    R23 = -max(3, int(round(151 * X / 2700)))
    C28 = max(3, int(round(200 * Y / 2100)))
    C38 = max(C28 + 2, int(round(250 * Y / 2100)))
    C_w = np.zeros(X)
    C_w[X - max(1, int(round(50 * X / 2700))):] = 1e-3
    return R23, C28, C38, C_w


def rt_densities(R, C, r0, r1):
    """init_rho_cosine (mrtcg_rayleigh_taylor.cpp:182-210) for global rows [r0, r1)"""
    s = R / 2.0 - 0.1 * C * np.cos(2.0 * 3.141592 * np.arange(C) / C)
    below = np.arange(r0, r1)[:, None] < s[None, :]
    return RED["rho_0"] * below.astype(np.float64), BLUE["rho_0"] * (~below).astype(np.float64)


def droplet_densities(Ln, radius, r0, r1):
    """init_rho (rk_static_droplet_test.cpp:363-396) for global rows [r0, r1)"""
    c = Ln / 2.0
    s = np.sqrt((np.arange(r0, r1)[:, None] - c) ** 2 + (np.arange(Ln)[None, :] - c) ** 2)
    sg = 1.0 / (1.0 + np.exp(-2.0 * (s - radius)))
    return RK_RED["rho_0"] * (1.0 - sg), RK_BLUE["rho_0"] * sg


def workload_config(args, n):
    wl = WORKLOADS[args.workload]
    gb = args.X * args.Y * 72.0 * wl["nlat"] / 1e9
    return {
        "workload": f"{wl['what']}, {args.X}x{args.Y} nodes per GPU",
        "grid_per_gpu": [args.X, args.Y], "global_grid": [args.X * n, args.Y], "decomposition": f"{n} slab(s) along axis 0",
        "l2": f"inputs larger than L2 (2 x {gb:.1f} GB of populations per GPU), no flush needed" if gb > 0.5 else
              f"populations per GPU: 2 x {gb * 1e3:.0f} MB (fits the 126 MB L2 when below that: reported as such)",
        "reference_driver": wl["driver"],
    }


# ---------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------
def cpu_reference_mlups(workload, edge, warmup, steps):
    """The reference's own CPU implementation (oracle/_ref: unmodified sources on CPU libtorch, all host
    threads) where the harness exposes the workload's loop (cylinder); the plain-C oracle port otherwise.
    Returns (MLUPS, seconds per step, kind, cores, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    omega, u_lb = lattice_parameters()
    port_cores = os.cpu_count() or 1  # the port's loops are OpenMP-parallel (oracle/Makefile: -fopenmp when available)

    def timed(step_fn):
        for _ in range(warmup):
            step_fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_fn()
        return (time.perf_counter() - t0) / max(steps, 1)

    if workload == "cylinder":
        xs, ys = cylinder_markers(edge, edge)
        desc = f"{edge}x{edge} crop of the cylinder workload"
        if oracle_lib.have_ref():
            ref = oracle_lib.Ref()
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "boundary.toml")
                write_markers_toml(path, xs, ys)
                sec, _ = ref.cylinder_loop(edge, edge, omega, u_lb, path, warmup, steps)
            return edge * edge / sec / 1e6, sec, "reference", ref.num_threads(), desc
        orc = oracle_lib.Oracle()
        ib = orc.ibm_create(xs, ys)
        u = np.zeros((edge, edge, 2)); u[..., 0] = u_lb
        rho = np.ones((edge, edge, 1))
        f = orc.incomp_equilibrium(u, rho)
        sec = timed(lambda: orc.cylinder_step(f, u, rho, omega, u_lb, ib))
        orc.ibm_destroy(ib)
        return edge * edge / sec / 1e6, sec, "port", port_cores, desc

    if workload == "cylinder_bb":
        # collide + stream through the reference's own solver:: operators (oracle/_ref) where they compiled, the port
        # otherwise; the bounce-back links are the driver-style slice assignments, done here with a mask per direction
        have = oracle_lib.have_ref()
        ops = oracle_lib.Ref() if have else oracle_lib.Oracle()
        solid = cylinder_solid(edge, edge, edge)
        opp = (0, 3, 4, 1, 2, 7, 8, 5, 6)
        cuts = [(q, solid != np.roll(solid, (int(CX9[q]), int(CY9[q])), axis=(0, 1))) for q in range(1, 9)]
        u = np.zeros((edge, edge, 2)); u[..., 0] = u_lb
        state = [ops.incomp_equilibrium(u, np.ones((edge, edge, 1)))]

        def bb_step():
            f = state[0]
            rho = ops.calc_rho(f)
            coll = ops.collision(f, ops.equilibrium(ops.calc_u(f, rho), rho), omega)
            adve = ops.advect(coll)
            for q, cut in cuts:
                adve[cut, q] = coll[cut, opp[q]]
            state[0] = adve

        sec = timed(bb_step)
        return (edge * edge / sec / 1e6, sec, "reference" if have else "port", ops.num_threads() if have else port_cores,
                f"{edge}x{edge} crop of the cylinder_bb workload (solver:: operators + bounce-back links; inlet/outlet rows left out)")

    if workload == "kbc_shear" and oracle_lib.have_ref():
        # the reference's own ulbm::d2q9::kbc object stepped like test/ulbm_double_shear_flow.cpp does (oracle/ref_harness.cpp)
        ref = oracle_lib.Ref()
        r = np.arange(edge)[:, None] + 0.0 * np.arange(edge)[None, :]
        c = np.arange(edge)[None, :] + 0.0 * np.arange(edge)[:, None]
        u = np.zeros((edge, edge, 2))
        u[..., 0] = 0.02 * np.tanh(80.0 * (0.25 * edge - np.abs(c - 0.5 * edge)))
        u[..., 1] = 0.02 * 0.05 * np.sin(6.2832 * (r + 0.25 * edge) / edge)
        m0 = np.ones((edge, edge))
        f = ref.kbc_equilibrium(m0, u)
        s2 = 1.0 / (0.5 + 3.0 * 1.70766666e-4)
        ref.kbc_run(f, m0, u, s2, max(warmup, 1))
        t0 = time.perf_counter()
        ref.kbc_run(f, m0, u, s2, steps)
        sec = (time.perf_counter() - t0) / max(steps, 1)
        return edge * edge / sec / 1e6, sec, "reference", ref.num_threads(), f"{edge}x{edge} crop of the kbc_shear workload"

    if workload == "poiseuille" and oracle_lib.have_ref() and hasattr(oracle_lib.Ref().lib, "ref_poiseuille_loop"):
        ref = oracle_lib.Ref()
        om, rho_in, rho_out = channel_constants(edge, edge)
        sec, _ = ref.poiseuille_loop(edge, edge, om, rho_in, rho_out, max(warmup, 1), steps)
        return edge * edge / sec / 1e6, sec, "reference", ref.num_threads(), f"{edge}x{edge} crop of the poiseuille workload"

    orc = oracle_lib.Oracle()
    if workload == "poiseuille":
        om, rho_in, rho_out = channel_constants(edge, edge)
        u = np.zeros((edge, edge, 2)); rho = np.ones((edge, edge, 1))
        f = orc.incomp_equilibrium(u, rho)
        sec = timed(lambda: orc.poiseuille_step(f, u, rho, om, rho_in, rho_out))
    elif workload in ("sedimentation", "sedimentation_ibm"):
        R23, C28, C38, C_w = sedimentation_geometry(edge, edge)
        f, g, u, rho, Cc = orc.sedimentation_init(edge, edge, u_lb, C_w)
        ib = orc.ibm_create(*sedimentation_body(edge, edge)) if workload == "sedimentation_ibm" else None
        sec = timed(lambda: orc.sedimentation_step(f, g, u, rho, Cc, omega, u_lb, 3e-3, C_w, R23, C28, C38, ib=ib))
        if ib is not None:
            orc.ibm_destroy(ib)
    elif workload == "mrtcg_rt":
        p = oracle_lib.MrtcgParams()
        p.R, p.C = edge, edge
        p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = RED["rho_0"], RED["alpha"], RED["nu"], RED["beta"]
        p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = BLUE["rho_0"], BLUE["alpha"], BLUE["nu"], BLUE["beta"]
        p.sigma, p.delta = 0.1, 0.1
        p.Fg[0], p.Fg[1] = RT_FG
        p.add_force = 1
        st = orc.mrtcg_init(p, "rt")
        sec = timed(lambda: orc.mrtcg_step(p, st))
    elif workload == "csf_rt":
        p = oracle_lib.CsfParams()
        p.R, p.C = edge, edge
        p.r_rho0, p.r_alpha, p.r_nu, p.r_beta, p.r_A = RED["rho_0"], RED["alpha"], RED["nu"], RED["beta"], RED["A"]
        p.b_rho0, p.b_alpha, p.b_nu, p.b_beta, p.b_A = BLUE["rho_0"], BLUE["alpha"], BLUE["nu"], BLUE["beta"], BLUE["A"]
        p.sigma, p.delta = 0.1, 0.1
        p.Fg[0], p.Fg[1] = RT_FG
        st = orc.csf_init(p)
        sec = timed(lambda: orc.csf_step(p, st))
    elif workload == "kbc_shear":
        r = np.arange(edge)[:, None] + 0.0 * np.arange(edge)[None, :]
        c = np.arange(edge)[None, :] + 0.0 * np.arange(edge)[:, None]
        u = np.zeros((edge, edge, 2))
        u[..., 0] = 0.02 * np.tanh(80.0 * (0.25 * edge - np.abs(c - 0.5 * edge)))
        u[..., 1] = 0.02 * 0.05 * np.sin(6.2832 * (r + 0.25 * edge) / edge)
        m0 = np.ones((edge, edge))
        f = orc.kbc_equilibrium(m0, u)
        s2 = 1.0 / (0.5 + 3.0 * 1.70766666e-4)
        sec = timed(lambda: orc.kbc_step(f, m0, u, s2))
    else:  # rk_droplet
        p = oracle_lib.RkParams()
        p.L, p.radius = edge, edge / 4.0
        p.r_rho0, p.r_alpha, p.r_A, p.r_nu = RK_RED["rho_0"], RK_RED["alpha"], RK_RED["A"], RK_RED["nu"]
        p.b_rho0, p.b_alpha, p.b_A, p.b_nu = RK_BLUE["rho_0"], RK_BLUE["alpha"], RK_BLUE["A"], RK_BLUE["nu"]
        p.delta = 0.98
        st = orc.rk_init(p)
        sec = timed(lambda: orc.rk_step(p, st))
    return edge * edge / sec / 1e6, sec, "port", port_cores, f"{edge}x{edge} crop of the {workload} workload"


def cpu_baseline_leg(workload, edge, target_seconds=12.0):
    """The `cpu_baseline` object of the B200 arm's line: the CPU implementation on a crop of the workload, sized from a short
    calibration run to about `target_seconds` of CPU work (the contract asks for 10-30 s; between 5 and 400 steps)."""
    _, sec, _, _, _ = cpu_reference_mlups(workload, edge, 1, 2)
    steps = int(min(400, max(5, round(target_seconds / max(sec, 1e-6)))))
    mlups, sec, kind, cores, desc = cpu_reference_mlups(workload, edge, 1, steps)
    out = {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
           "sample": f"{desc}, {steps} steps after 1 warm-up ({sec * steps:.1f} s of CPU work)", "ms_per_step": sec * 1e3}
    if kind == "reference":
        # the reference also on ONE host thread (at::set_num_threads(1)), a few steps of the same crop
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib

        ref = oracle_lib.Ref()
        ref.lib.ref_set_num_threads(1)
        try:
            one, _, _, _, _ = cpu_reference_mlups(workload, edge, 1, int(min(steps, max(2, round(4.0 / max(sec * cores, 1e-6))))))
            out["single_thread"] = {"value": one, "unit": "MLUPS", "cores": 1}
        finally:
            ref.lib.ref_set_num_threads(cores)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    edge = args.cpu_sample
    mlups, sec, kind, cores, desc = cpu_reference_mlups(args.workload, edge, args.warmup, args.steps)
    sample = f"{desc}, {args.steps} steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up; the reported clocks are those of the samples whose
    timestamps fall inside the timed region (mark_begin / mark_end)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        import datetime

        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = time.time()
            self.rows.append([ts] + c[1:])

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        isnum = lambda v: v.replace(".", "").isdigit()
        rows = [r for r in self.rows if len(r) > 2 and isnum(r[1])]
        t0, t1 = self.t0 or 0.0, self.t1 or 1e30
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if not inside and rows:  # region shorter than the polling period: the sample nearest to it
            mid = 0.5 * (t0 + t1)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
        sm = [float(r[1]) for r in inside]
        mx = [float(r[2]) for r in inside if isnum(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in inside)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_total": len(rows)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(workload, X, Y):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (profiles/roofline_traffic.json), only when it was taken at this grid size; on the same basis as `achieved`
    (all of a step's launches of that kernel: the single-phase step has an early-rows and a bulk-rows launch)"""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get(f"{workload}_{X}x{Y}_bytes_per_launch")
    except Exception:
        return None


class Case:
    """One workload on one slab: builds the domain, holds the pinned host state, imports it."""

    def __init__(self, L, torch, args, rank, world, local):
        self.L, self.torch, self.args = L, torch, args
        self.rank, self.world, self.local = rank, world, local
        self.X, self.Y = args.X, args.Y
        self.Xg = args.X * world
        self.x0, self.x1 = rank * args.X, (rank + 1) * args.X
        self.name = args.workload
        self.omega, self.u_lb = lattice_parameters()
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        slab = dict(X=self.Xg, Y=self.Y, x0=self.x0, x1=self.x1, device=local)
        if self.name == "cylinder":
            cfg = L.default_config(model=L.MODEL_BGK, omega=self.omega, equilibrium=L.EQ_COMPRESSIBLE, force=L.FORCE_IBM, **slab)
        elif self.name == "cylinder_bb":
            cfg = L.default_config(model=L.MODEL_BGK, omega=self.omega, equilibrium=L.EQ_COMPRESSIBLE, **slab)
        elif self.name == "poiseuille":
            om, self.rho_in, self.rho_out = channel_constants(self.Xg, self.Y)
            cfg = L.default_config(model=L.MODEL_BGK, omega=om, equilibrium=L.EQ_INCOMPRESSIBLE, **slab)
        elif self.name == "kbc_shear":
            cfg = L.default_config(model=L.MODEL_KBC, omega=1.0 / (0.5 + 3.0 * 1.70766666e-4), **slab)
        elif self.name in ("sedimentation", "sedimentation_ibm"):
            cfg = L.default_config(model=L.MODEL_BGK_ADE, omega=self.omega, omega_g=self.omega, equilibrium=L.EQ_COMPRESSIBLE,
                                   w_s=3e-3, force=L.FORCE_IBM if self.name == "sedimentation_ibm" else L.FORCE_NONE, **slab)
        elif self.name == "csf_rt":
            cfg = L.default_config(model=L.MODEL_MRT_CSF, red=RED, blue=BLUE, sigma=0.1, delta=0.1, Fg=RT_FG, add_force=1, **slab)
        elif self.name == "mrtcg_rt":
            cfg = L.default_config(model=L.MODEL_MRTCG, red=RED, blue=BLUE, sigma=0.1, delta=0.1, Fg=RT_FG, add_force=1, **slab)
        else:
            cfg = L.default_config(model=L.MODEL_RK, red=RK_RED, blue=RK_BLUE, delta=0.98, **slab)
        self.d = L.Domain(cfg)

    def comm_init(self, ident):
        if self.world > 1:
            self.d.comm_init(ident, self.world, self.rank)

    def pinned(self, shape):
        t = self.torch.empty(shape, dtype=self.torch.float64, pin_memory=True)
        return t, t.numpy()

    def setup(self):
        d, X, Y = self.d, self.X, self.Y
        if self.name in ("cylinder", "cylinder_bb"):
            d.preset_free_stream(self.u_lb, 0.0)
            if self.name == "cylinder_bb":
                d.bc_add_solid(cylinder_solid(self.Xg, Y, self.X))  # appended to the preset's rules; the body sits in rank 0's slab
                d.bc_commit()
            else:
                xs, ys = cylinder_markers(self.X, Y)  # the body sits in rank 0's slab
                if self.rank == 0:
                    d.ibm_set_markers(xs, ys)
            # the drivers' f = incomp_equilibrium(u=(u_lb,0), rho=1) (cylinder_test.cpp:84-86): u and rho are the inputs
            self.ri_t, self.ri = self.pinned((X, Y, 1))
            self.ui_t, self.ui = self.pinned((X, Y, 2))
            self.ri[...] = 1.0
            self.ui[..., 0] = self.u_lb
            self.ui[..., 1] = 0.0
        elif self.name == "poiseuille":
            d.preset_poiseuille(self.rho_in, self.rho_out)
            self.ri_t, self.ri = self.pinned((X, Y, 1))  # incomp_equilibrium(u=0, rho=1) (horizontal_poiseuille_test.cpp:91)
            self.ui_t, self.ui = self.pinned((X, Y, 2))
            self.ri[...] = 1.0
            self.ui[...] = 0.0
        elif self.name == "kbc_shear":
            d.preset_periodic()
            self.ri_t, self.ri = self.pinned((X, Y, 1))
            self.ui_t, self.ui = self.pinned((X, Y, 2))
            self.ri[...] = 1.0   # set_initial_conditions (ulbm_double_shear_flow.cpp:44-67) on global rows [x0, x1)
            r = np.arange(self.x0, self.x1)[:, None]; c = np.arange(Y)[None, :]
            self.ui[..., 0] = 0.02 * np.tanh(80.0 * (0.25 * self.Xg - np.abs(c - 0.5 * self.Xg))) + 0.0 * r
            self.ui[..., 1] = 0.02 * 0.05 * np.sin(6.2832 * (r + 0.25 * self.Xg) / self.Xg) + 0.0 * c
        elif self.name in ("sedimentation", "sedimentation_ibm"):
            R23, C28, C38, C_w = sedimentation_geometry(self.Xg, Y)
            d.preset_sedimentation(self.u_lb, C_w, R23, C28, C38)
            if self.name == "sedimentation_ibm" and self.rank == 0:
                d.ibm_set_markers(*sedimentation_body(self.X, Y))  # the body sits in rank 0's slab
            # rectangle_sedimentation_test.cpp:84-103: u = (0, u_lb); f = incomp_eq(u, 1); g = eq(u, C), C = C_w on column 0
            self.f_t, self.f = self.pinned((X, Y, 9))
            self.g_t, self.g = self.pinned((X, Y, 9))
            self.f[...] = (1.0 + 3.0 * CY9 * self.u_lb) * W9
            cu = CY9 * self.u_lb
            geq = W9 * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * self.u_lb ** 2)
            self.g[...] = 0.0
            self.g[:, 0, :] = C_w[self.x0:self.x1, None] * geq[None, :]
        else:
            (d.preset_rk if self.name == "rk_droplet" else d.preset_mrtcg)()
            self.rr_t, self.rr = self.pinned((X, Y))
            self.rb_t, self.rb = self.pinned((X, Y))
            self.u_t, self.u = self.pinned((X, Y, 2))
            self.u[...] = 0.0
            if self.name != "rk_droplet":
                self.rr[...], self.rb[...] = rt_densities(self.Xg, Y, self.x0, self.x1)
            else:
                self.rr[...], self.rb[...] = droplet_densities(self.Xg, self.Xg / 4.0, self.x0, self.x1)
        self.rho_t, self.rho = self.pinned((X, Y, 1))
        self.uo_t, self.uo = self.pinned((X, Y, 2))

    def import_state(self):
        """host -> device through the C ABI; returns the bytes copied"""
        d = self.d
        if self.name in ("cylinder", "cylinder_bb", "poiseuille"):
            d.init_equilibrium(self.ri, self.ui, self.L.EQ_INCOMPRESSIBLE)
            return self.ri.nbytes + self.ui.nbytes
        if self.name == "kbc_shear":
            d.init_equilibrium(self.ri, self.ui, self.L.EQ_KBC_FRESH)
            d.set_moments(self.ri, self.ui)
            return 2 * (self.ri.nbytes + self.ui.nbytes)
        if self.name in ("sedimentation", "sedimentation_ibm"):
            d.set_f(self.f, 0)
            d.set_f(self.g, 1)
            return self.f.nbytes + self.g.nbytes
        d.init_two_phase(self.rr, self.rb, self.u)
        return self.rr.nbytes + self.rb.nbytes + self.u.nbytes

    def export_moments(self):
        """device -> host: rho and u of the current state (what every driver snapshots)"""
        lib = self.L.load()
        dp = ctypes.POINTER(ctypes.c_double)
        rc = lib.lbm_get_moments(self.d.h, 0, self.rho.ctypes.data_as(dp), self.uo.ctypes.data_as(dp))
        if rc != 0:
            raise RuntimeError(lib.lbm_last_error().decode())
        return self.rho.nbytes + self.uo.nbytes

    def dominant_classes(self):
        L = self.L
        return [L.PROF_INTERIOR]

    def dominant_nodes(self):
        """nodes per step the dominant kernel owns (the remaining edge columns are listed nodes)"""
        if self.name in ("mrtcg_rt", "rk_droplet", "csf_rt"):
            return self.X * (self.Y - 2)
        return self.X * (2 * ((self.Y - 3) // 2))


def run_b200_arm(args):
    import torch

    import lbm_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = WORKLOADS[args.workload]
    case = Case(L, torch, args, rank, world, local)
    d = case.d
    if world > 1:
        ident = [L.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        case.comm_init(ident[0])
    case.setup()

    def barrier():
        d.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    case.import_state()
    d.step(args.warmup)
    barrier()

    # ---- device-resident timed region: exactly K steps, CUDA events on the domain's stream
    launches0 = d.kernel_launches()
    d.profile_enable(True)
    barrier()
    sampler.mark_begin()
    d.step(args.steps)
    d.synchronize()
    ms = d.last_step_ms()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = d.kernel_launches() - launches0
    prof = {name: d.profile_read(cls) for name, cls in
            [("interior", L.PROF_INTERIOR), ("boundary", L.PROF_BOUNDARY), ("fixup", L.PROF_FIXUP),
             ("ghost", L.PROF_GHOST), ("ibm", L.PROF_IBM), ("moments", L.PROF_MOMENTS)]}
    dom_ms = sum(d.profile_read(c)[0] for c in case.dominant_classes())
    dom_n = sum(d.profile_read(c)[1] for c in case.dominant_classes())
    d.profile_enable(False)
    ms = max_over_ranks(ms)
    mlups = (case.Xg * case.Y) * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers: import the state (H2D), K steps, export rho,u (D2H)
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        h2d = case.import_state()
        t1 = time.perf_counter()
        d.step(args.steps)
        d2h = case.export_moments()
        d.synchronize()
        t2 = time.perf_counter()
        sec = max_over_ranks(t2 - t0)
        e2e = {"value": (case.Xg * case.Y) * args.steps / sec / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "what": f"initial fields from pinned host (the drivers' u, rho -> equilibrium; populations for the ADE / two-phase "
                       f"imports) + lbm_step({args.steps}) + lbm_get_moments to pinned host, per rank",
               "seconds": sec, "import_seconds": t1 - t0, "steps_and_export_seconds": t2 - t1,
               "rho_mean": float(case.rho_t.mean())}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, timed with CUDA events on its own stream inside the timed region above
    peak, peak_src = measured_hbm_peak()
    B = wl["bytes"]
    nodes_per_step = case.dominant_nodes()
    achieved = B * nodes_per_step * args.steps / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None
    per_gpu_mlups = mlups / world
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None,
                "traffic": ncu_traffic_per_launch(args.workload, args.X, args.Y),
                "kernel": ("k_csf_fused (LBM_CSF_FUSED=1: one pass)" if args.workload == "csf_rt" and os.environ.get("LBM_CSF_FUSED") == "1"
                           else wl["kernel"]), "bytes_per_node": B,
                "algorithmic_bytes_per_step": B * nodes_per_step,
                "launches_per_step": dom_n / args.steps if args.steps else 0, "kernel_ms_per_step": dom_ms / args.steps,
                "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0 if achieved else None,
                "whole_step_frac_per_gpu": B * per_gpu_mlups * 1e6 / 1e9 / peak,
                "whole_step_frac_per_gpu_of_nominal_8TBs": B * per_gpu_mlups * 1e6 / 1e9 / 8000.0,
                "share_of_step": dom_ms / ms,
                "other_spans_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()
                                            if v[1] and k != "interior"},
                "note": f"achieved = {B:.0f} B x nodes the dominant kernel owns per step / summed duration of its launches "
                        "(rank 0); listed nodes, stages, ghost rows and the IBM pre-pass run on a side stream under the bulk launch"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(args.workload, args.cpu_sample)

    line = {
        "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
