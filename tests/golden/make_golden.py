#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs the reference's own driver binaries (oracle/_ref/<driver>, built by `make -C oracle ref`
from the sources under /root/reference — CPU libtorch, torch::kCUDA re-pointed at kCPU) and
stores small slices of what they write with torch::save.  While doing so it asserts that the
plain-C oracle (oracle/lbm_oracle.c) reproduces every stored number, which is what pins the
oracle.  Needs /root/reference + oracle/_ref; the committed .npz files are what the GPU box uses.

usage: python tests/golden/make_golden.py [case ...]   (default: all cases)
"""
import ctypes as C
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import CsfParams, MrtcgParams, Oracle, RkParams, load_pt, run_ref_driver  # noqa: E402

WORK = os.environ.get("GOLDEN_WORK", "/tmp/refrun")
ORC = Oracle()


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def workdir(name):
    d = os.path.join(WORK, name)
    os.makedirs(d, exist_ok=True)
    return d


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.0f} KiB)")


def channel_constants(H, W, u_max):
    tau = np.sqrt(3.0 / 16.0) + 0.5
    omega = 1.0 / tau
    nu = (2.0 * tau - 1.0) / 6.0
    p_grad = 8.0 * nu * u_max / (W * W)
    rho_out = 1.0
    rho_in = 3.0 * (H - 1) * p_grad + rho_out
    return omega, rho_in, rho_out


# ------------------------------------------------------------------ driver 10
def case_poiseuille():
    d = workdir("hpt")
    if not os.path.exists(os.path.join(d, "hpt-fs.pt")):
        r = run_ref_driver("horizontal_poiseuille_test", [], d)
        assert r.returncode == 0, r.stderr
        open(os.path.join(d, "stdout.txt"), "w").write(r.stdout)
    out = open(os.path.join(d, "stdout.txt")).read() if os.path.exists(os.path.join(d, "stdout.txt")) else ""
    fs = load_pt(os.path.join(d, "hpt-fs.pt"))  # {H,W,9,T}; fs[...,t] = f_adve at the START of iteration t
    ux = load_pt(os.path.join(d, "hpt-ux.pt"))
    uy = load_pt(os.path.join(d, "hpt-uy.pt"))
    ps = load_pt(os.path.join(d, "hpt-ps.pt"))
    H, W, _, T = fs.shape
    omega, rho_in, rho_out = channel_constants(H, W, 1.030985714e-1)
    f = np.ascontiguousarray(fs[..., 0]).copy()
    u = np.zeros((H, W, 2)); rho = np.ones((H, W, 1))
    keep = [0, 1, 2, 3, 10, 100, 1000, 5000, T - 1]
    worst = 0.0
    for t in range(T - 1):
        ORC.poiseuille_step(f, u, rho, omega, rho_in, rho_out)
        e = relerr(f, fs[..., t + 1])
        worst = max(worst, e)
        # u/rho saved at iteration t+1 are the moments computed in iteration t
        worst = max(worst, float(np.abs(u[..., 0] - ux[..., t + 1]).max()), float(np.abs(rho[..., 0] / 3.0 - ps[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {T - 1} steps: worst rel err {worst:.3e}")
    assert worst < 1e-12
    l2 = None
    for line in out.splitlines():
        if line.startswith("L2="):
            l2 = float(line[3:])
    save("poiseuille_21x21", steps=np.array(keep), f=np.stack([fs[..., t] for t in keep]),
         ux=np.stack([ux[..., t] for t in keep]), uy=np.stack([uy[..., t] for t in keep]),
         ps=np.stack([ps[..., t] for t in keep]), omega=omega, rho_in=rho_in, rho_out=rho_out,
         l2=np.array(l2 if l2 is not None else np.nan))


# ------------------------------------------------------------------ driver 13
def case_specular():
    d = workdir("sbt")
    if not os.path.exists(os.path.join(d, "sbt-fs.pt")):
        r = run_ref_driver("specular_boundary_test", [], d)
        assert r.returncode == 0, r.stderr
    fs = load_pt(os.path.join(d, "sbt-fs.pt"))
    ux = load_pt(os.path.join(d, "sbt-ux.pt"))
    H, W, _, T = fs.shape
    omega, rho_in, rho_out = channel_constants(H, W, 0.1)
    f = np.ascontiguousarray(fs[..., 0]).copy()
    u = np.zeros((H, W, 2)); rho = np.ones((H, W, 1))
    NS = 2000
    keep = [0, 1, 2, 10, 100, 1000, NS]
    worst = 0.0
    for t in range(NS):
        ORC.specular_step(f, u, rho, omega, rho_in, rho_out)
        worst = max(worst, relerr(f, fs[..., t + 1]), float(np.abs(u[..., 0] - ux[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {NS} steps: worst rel err {worst:.3e}")
    assert worst < 1e-12
    save("specular_51x51", steps=np.array(keep), f=np.stack([fs[..., t] for t in keep]),
         omega=omega, rho_in=rho_in, rho_out=rho_out)


# ------------------------------------------------------------------ driver 14
def case_gravity():
    d = workdir("gt")
    if not os.path.exists(os.path.join(d, "gt-fs.pt")):
        r = run_ref_driver("gravity_test", [], d)
        assert r.returncode == 0, r.stderr
    fs = load_pt(os.path.join(d, "gt-fs.pt"))
    ux = load_pt(os.path.join(d, "gt-ux.pt"))
    H, W, _, T = fs.shape
    omega, _, _ = channel_constants(H, W, 0.1)
    rho_in = rho_out = 1.0
    Fg = np.array([-0.0003, 0.0])
    f = np.ascontiguousarray(fs[..., 0]).copy()
    u = np.zeros((H, W, 2)); rho = np.ones((H, W, 1))
    NS = 3000
    keep = [0, 1, 2, 10, 100, 1000, NS]
    worst = 0.0
    for t in range(NS):
        ORC.gravity_step(f, u, rho, omega, rho_in, rho_out, Fg)
        worst = max(worst, relerr(f, fs[..., t + 1]), float(np.abs(u[..., 0] - ux[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {NS} steps: worst rel err {worst:.3e}")
    assert worst < 1e-12
    save("gravity_21x21", steps=np.array(keep), f=np.stack([fs[..., t] for t in keep]),
         ux=np.stack([ux[..., t] for t in keep]), omega=omega, rho_in=rho_in, rho_out=rho_out, Fg=Fg)


# ------------------------------------------------------------------ driver 19
def case_decompose():
    d = workdir("dd")
    if not os.path.exists(os.path.join(d, "B-domain-decomp-hpt-fs.pt")):
        r = run_ref_driver("decompose_domain", [], d)
        assert r.returncode == 0, r.stderr
    A = load_pt(os.path.join(d, "A-domain-decomp-hpt-fs.pt"))
    B = load_pt(os.path.join(d, "B-domain-decomp-hpt-fs.pt"))
    H, W, _, T = A.shape
    omega, rho_in, rho_out = channel_constants(H, W, 1.030985714e-1)
    fA = np.ascontiguousarray(A[..., 0]).copy(); fB = np.ascontiguousarray(B[..., 0]).copy()
    uA = np.zeros((H, W, 2)); uB = np.zeros((H, W, 2)); rA = np.ones((H, W, 1)); rB = np.ones((H, W, 1))
    keep = [0, 1, 2, 10, 100, T - 1]
    worst = 0.0
    for t in range(T - 1):
        ORC.decompose_step(fA, uA, rA, fB, uB, rB, omega, rho_in, rho_out)
        worst = max(worst, relerr(fA, A[..., t + 1]), relerr(fB, B[..., t + 1]))
    print(f"  oracle vs reference driver over {T - 1} steps: worst rel err {worst:.3e}")
    assert worst < 1e-12
    save("decompose_2x21x21", steps=np.array(keep), fA=np.stack([A[..., t] for t in keep]),
         fB=np.stack([B[..., t] for t in keep]), omega=omega, rho_in=rho_in, rho_out=rho_out)


def case_loop_blocks():
    """test/decompose_domain_loop.cpp with L = 128, T = 400 (oracle/Makefile: the same translation unit, two constants
    lowered): m_1, m_0 of the four blocks every 50 iterations.  Pins oracle/blocks_oracle.py."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import blocks_oracle as BO
    d = workdir("loop")
    if not os.path.exists(os.path.join(d, "D-domain-decomp-hpt-rho.pt")):
        r = run_ref_driver("decompose_domain_loop_small", [], d)
        assert r.returncode == 0, r.stderr
    Ln = 128
    omega = 1.0 / (np.sqrt(3.0 / 16.0) + 0.5)
    ref = {k: {f: load_pt(os.path.join(d, f"{k}-domain-decomp-hpt-{f}.pt")) for f in ("ux", "uy", "rho")} for k in "ABCD"}
    T = ref["A"]["ux"].shape[-1]
    st = BO.init(ORC, Ln)
    worst = 0.0
    for t in range(50 * (T - 1) + 1):
        if t % 50 == 0:   # the snapshot at the top of iteration t: m_0, m_1 of iteration t - 1, with F added on A's forced rows
            ts = t // 50
            for k in "ABCD":
                u = st[k]["u"].copy()
                if k == "A":
                    u[BO.force_rows(Ln), :, 0] += 3e-3
                worst = max(worst, np.abs(u[..., 0] - ref[k]["ux"][..., ts]).max(), np.abs(u[..., 1] - ref[k]["uy"][..., ts]).max(),
                            np.abs(st[k]["rho"][..., 0] - ref[k]["rho"][..., ts]).max())
        if t == 0:
            # the driver adds F to m_1 at the top of EVERY iteration and only iteration t's calc_u overwrites it (:116,143):
            # at t = 0 nothing has been computed yet, m_1 = F on the forced rows, which the snapshot shows and calc_u discards
            pass
        BO.step(ORC, st, Ln, omega)
    print(f"  oracle vs reference driver over {T} snapshots (L = {Ln}): worst abs err {worst:.3e}")
    assert worst < 1e-12
    out = {}
    keep = [0, 1, 2, T - 1]
    for k in "ABCD":
        for f in ("ux", "uy", "rho"):
            out[f"{k}_{f}"] = np.moveaxis(ref[k][f], -1, 0)[keep]
    save("loop_blocks_128", L=Ln, omega=omega, steps=np.array(keep) * 50, **out)


# ------------------------------------------------------------------ TOML-driven drivers
PARAMS_TOML = """\
[flow]
initial_density = 1e3
kinematic_viscosity = 1.0e-3
characteristic_length = 0.011
characteristic_velocity = {U}

[lattice]
relaxation_time = {TAU}
lattice_spacing = 1.0e-3
x_multiplier = {XM}
y_multiplier = {YM}

[simulation]
stop_time = {STOP}
snapshot_period = 0.00005
file_prefix = "{PREFIX}"
"""


def write_params(path, U, TAU, XM, YM, nsteps, prefix):
    # dt = cs2 (tau-0.5) dx^2 / nu ; T = ceil(1/dt); total_steps = ceil(stop*T)
    dt = (1.0 / 3.0) * (TAU - 0.5) * (1.0e-3 * 1.0e-3) / 1.0e-3
    T = int(np.ceil(1.0 / dt))
    stop = (nsteps - 0.5) / T
    open(path, "w").write(PARAMS_TOML.format(U=U, TAU=TAU, XM=XM, YM=YM, STOP=repr(stop), PREFIX=prefix))
    return T


def ref_lattice(path, with_simulation=True):
    from oracle_lib import Ref

    return Ref().params(path, with_simulation)


# ------------------------------------------------------------------ driver 12
def case_free_stream():
    d = workdir("fst")
    toml = os.path.join(d, "params.toml")
    NS = 60
    write_params(toml, U=0.5, TAU=0.8, XM=3, YM=2, nsteps=NS, prefix="g-")
    lp = ref_lattice(toml)
    assert int(lp["total_steps"]) == NS and int(lp["snapshot_steps"]) == 1, lp
    r = run_ref_driver("free_stream_test", [toml, "go"], d)
    assert r.returncode == 0, r.stderr
    ux = load_pt(os.path.join(d, "g-fst-ux.pt")); uy = load_pt(os.path.join(d, "g-fst-uy.pt"))
    ps = load_pt(os.path.join(d, "g-fst-ps.pt"))
    X, Y = int(lp["X"]), int(lp["Y"])
    omega = lp["omega"]
    u = np.zeros((X, Y, 2)); u[..., 0] = 0.1
    rho = np.ones((X, Y, 1))
    f = ORC.incomp_equilibrium(u, rho)
    f0 = f.copy()
    worst = 0.0
    for t in range(NS - 1):
        ORC.free_stream_step(f, u, rho, omega, 0.1)
        # snapshot t+1 holds the moments computed during iteration t
        worst = max(worst, float(np.abs(u[..., 0] - ux[..., t + 1]).max()), float(np.abs(u[..., 1] - uy[..., t + 1]).max()),
                    float(np.abs(rho[..., 0] / 3.0 - ps[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {NS - 1} steps ({X}x{Y}): worst abs err {worst:.3e}")
    assert worst < 1e-13
    save("free_stream_33x22", ux=np.moveaxis(ux, -1, 0), uy=np.moveaxis(uy, -1, 0), ps=np.moveaxis(ps, -1, 0),
         omega=omega, X=X, Y=Y, f0=f0, uwx=0.1, toml=open(toml).read())


def circle_markers(cx, cy, r, n):
    th = 2.0 * np.pi * np.arange(n) / n
    return cx + r * np.cos(th), cy + r * np.sin(th)


def write_boundary(path, name, xs, ys):
    with open(path, "w") as fh:
        fh.write(f"[{name}]\n")
        fh.write("x = [" + ", ".join(repr(float(v)) for v in xs) + "]\n")
        fh.write("y = [\n" + ",\n".join("  " + repr(float(v)) for v in ys) + "\n]\n")


# ------------------------------------------------------------------ driver 11 (+ src/ibm.cpp)
def case_cylinder():
    from oracle_lib import Ref

    d = workdir("ct")
    toml = os.path.join(d, "params.toml")
    btoml = os.path.join(d, "boundary.toml")
    NS = 40
    write_params(toml, U=0.5, TAU=0.8, XM=9, YM=7, nsteps=NS, prefix="g-")
    lp = ref_lattice(toml)
    assert int(lp["total_steps"]) == NS and int(lp["snapshot_steps"]) == 1, lp
    X, Y = int(lp["X"]), int(lp["Y"])
    # the driver hard-codes a {45,44,2} force snapshot (cylinder_test.cpp:63) => the ROI must be 45x44
    xs, ys = circle_markers(50.5, 40.1, 19.8, 124)
    write_boundary(btoml, "cylinder-a", xs, ys)
    r = run_ref_driver("cylinder_test", [toml, btoml, "go"], d)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    ux = load_pt(os.path.join(d, "g-ct-ux.pt")); uy = load_pt(os.path.join(d, "g-ct-uy.pt"))
    ps = load_pt(os.path.join(d, "g-ct-ps.pt")); Fs = load_pt(os.path.join(d, "g-ct-Fs.pt"))
    FF = load_pt(os.path.join(d, "g-ct-F.pt"))
    omega, u_lb = lp["omega"], lp["u"]
    ib = ORC.ibm_create(xs, ys)
    roi = ORC.ibm_roi(ib)
    assert (roi[1] - roi[0], roi[3] - roi[2]) == (45, 44), roi
    # IBM alone, against the reference class on a random field
    rng = np.random.default_rng(7)
    ut = 0.05 * rng.standard_normal((X, Y, 2)); rt = 1.0 + 0.01 * rng.standard_normal((X, Y, 1))
    roi_r, F_r = Ref().ibm_force(btoml, "cylinder-a", ut, rt)
    assert tuple(roi_r) == tuple(roi)
    F_o = ORC.ibm_force(ib, ut, rt)
    e = relerr(F_o, F_r)
    print(f"  ibm force oracle vs reference class: rel err {e:.3e}")
    assert e < 1e-13
    u = np.zeros((X, Y, 2)); u[..., 0] = u_lb
    rho = np.ones((X, Y, 1))
    f = ORC.incomp_equilibrium(u, rho)
    f0 = f.copy()
    worst = 0.0
    for t in range(NS - 1):
        F = ORC.cylinder_step(f, u, rho, omega, u_lb, ib)
        worst = max(worst, float(np.abs(u[..., 0] - ux[..., t + 1]).max()), float(np.abs(u[..., 1] - uy[..., t + 1]).max()),
                    float(np.abs(rho[..., 0] / 3.0 - ps[..., t + 1]).max()), float(np.abs(F - FF[..., t + 1]).max()),
                    float(np.abs(F.reshape(-1, 2).sum(0) - Fs[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {NS - 1} steps ({X}x{Y}): worst abs err {worst:.3e}")
    assert worst < 1e-13
    ORC.ibm_destroy(ib)
    K = 16  # snapshots kept in the fixture (the comparison above covers all of them)
    save("cylinder_99x77", ux=np.moveaxis(ux, -1, 0)[:K], uy=np.moveaxis(uy, -1, 0)[:K], ps=np.moveaxis(ps, -1, 0)[:K],
         Fs=np.moveaxis(Fs, -1, 0)[:K], F=np.moveaxis(FF, -1, 0)[:K], omega=omega, u_lb=u_lb, X=X, Y=Y, f0=f0,
         marker_x=xs, marker_y=ys, roi=np.array(roi), ibm_u=ut, ibm_rho=rt, ibm_F=F_r,
         toml=open(toml).read())


# ------------------------------------------------------------------ driver 15
def case_sedimentation():
    d = workdir("rst")
    toml = os.path.join(d, "params.toml")
    NS = 30
    write_params(toml, U=0.25, TAU=0.8, XM=16, YM=24, nsteps=NS, prefix="g")
    lp = ref_lattice(toml)
    assert int(lp["total_steps"]) == NS and int(lp["snapshot_steps"]) == 1, lp
    X, Y = int(lp["X"]), int(lp["Y"])
    assert X > 160 and Y > 260
    r = run_ref_driver("rectangle_sedimentation_test", [toml], d)
    assert r.returncode == 0, r.stderr
    ux = load_pt(os.path.join(d, "g-ux.pt")); uy = load_pt(os.path.join(d, "g-uy.pt"))
    ps = load_pt(os.path.join(d, "g-ps.pt")); cs = load_pt(os.path.join(d, "g-cs.pt"))
    omega, u_lb = lp["omega"], lp["u"]
    C_w = np.zeros(X); C_w[-50:] = 1e-3
    f, g, u, rho, Cc = ORC.sedimentation_init(X, Y, u_lb, C_w)
    f0, g0 = f.copy(), g.copy()
    worst = max(float(np.abs(u[..., 1] - uy[..., 0]).max()), float(np.abs(Cc[..., 0] - cs[..., 0]).max()))
    for t in range(NS - 1):
        ORC.sedimentation_step(f, g, u, rho, Cc, omega, u_lb, 3e-3, C_w, -151, 200, 250)
        worst = max(worst, float(np.abs(u[..., 0] - ux[..., t + 1]).max()), float(np.abs(u[..., 1] - uy[..., t + 1]).max()),
                    float(np.abs(rho[..., 0] / 3.0 - ps[..., t + 1]).max()), float(np.abs(Cc[..., 0] - cs[..., t + 1]).max()))
    print(f"  oracle vs reference driver over {NS - 1} steps ({X}x{Y}): worst abs err {worst:.3e}")
    assert worst < 1e-13
    keep = [0, 1, 2, 5, 10, NS - 1]
    save("sedimentation_176x264", steps=np.array(keep), ux=np.stack([ux[..., t] for t in keep]),
         uy=np.stack([uy[..., t] for t in keep]), ps=np.stack([ps[..., t] for t in keep]),
         cs=np.stack([cs[..., t] for t in keep]), omega=omega, u_lb=u_lb, X=X, Y=Y, w_s=3e-3, C_w=C_w,
         walls=np.array([-151, 200, 250]), toml=open(toml).read())


MRTCG_TOML = """\
delta = 0.1
{GENERAL}
[domain]
rows = {R}
columns = {C}
time_steps = {T}
nr_snapshots = {T}

[red]
initial_density = 3.0
alpha = 0.7
kinematic_viscosity = 0.04
interfacial_tension_control = 0.5 # A
interface_thickness_control = 0.7 # beta

[blue]
initial_density = 1.0
alpha = 0.1
kinematic_viscosity = 0.04
interfacial_tension_control = 0.5 # A
interface_thickness_control = -0.7 # beta
"""


def mrtcg_params(R, Cc, sigma, Fg, add_force):
    p = MrtcgParams()
    p.R, p.C = R, Cc
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = 3.0, 0.7, 0.04, 0.7
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = 1.0, 0.1, 0.04, -0.7
    p.sigma, p.delta = sigma, 0.1
    p.Fg[0], p.Fg[1] = Fg
    p.add_force = add_force
    return p


def compare_mrtcg(prefix, d, p, kind, NS, extra=()):
    rhos = load_pt(os.path.join(d, prefix + "rhos.pt")); uxs = load_pt(os.path.join(d, prefix + "uxs.pt"))
    uys = load_pt(os.path.join(d, prefix + "uys.pt")); snus = load_pt(os.path.join(d, prefix + "snus.pt"))
    phases = load_pt(os.path.join(d, prefix + "phases.pt"))
    st = ORC.mrtcg_init(p, kind)
    init = {k: v.copy() for k, v in st.items()}
    worst = 0.0
    # snapshot t holds rho,u at the START of iteration t and s_nu/phase of iteration t-1
    for t in range(NS):
        worst = max(worst, float(np.abs(st["rho"][..., 0] - rhos[..., t]).max()),
                    float(np.abs(st["u"][..., 0] - uxs[..., t]).max()), float(np.abs(st["u"][..., 1] - uys[..., t]).max()),
                    float(np.abs(st["s_nu"] - snus[..., t]).max()), float(np.abs(st["phase"][..., 0] - phases[..., t]).max()))
        ORC.mrtcg_step(p, st)
    return worst, init, dict(rhos=rhos, uxs=uxs, uys=uys, snus=snus, phases=phases)


# ------------------------------------------------------------------ driver 16
def case_mrtcg_rt():
    d = workdir("mrtcg_rt")
    toml = os.path.join(d, "rt.toml")
    # 40 steps: with this sharp-interface start the reference's own scheme drives rho_blue slightly
    # negative on wall row 0 after ~50 steps and rounding differences then grow ~3x per step
    NS, R, Cc = 40, 64, 48
    open(toml, "w").write(MRTCG_TOML.format(
        GENERAL='\n[general]\nsigma = 0.1\ngravity_magnitude = 6.25e-6\nname = "g"\n', R=R, C=Cc, T=NS))
    r = run_ref_driver("mrtcg_rayleigh_taylor", [toml], d)
    assert r.returncode == 0, r.stderr
    p = mrtcg_params(R, Cc, 0.1, (6.25e-6, 0.0), 1)
    worst, init, ref = compare_mrtcg("g-mrtcg-rayleigh-taylor-", d, p, "rt", NS)
    print(f"  oracle vs reference driver over {NS} snapshots ({R}x{Cc}): worst abs err {worst:.3e}")
    assert worst < 1e-12
    keep = [0, 1, 2, 5, 10, 20, NS - 1]
    save("mrtcg_rt_64x48", steps=np.array(keep), **{k: np.stack([v[..., t] for t in keep]) for k, v in ref.items()},
         toml=open(toml).read())


# ------------------------------------------------------------------ driver 21 (CSF variant)
def csf_params(R, Cc):
    p = CsfParams()
    p.R, p.C = R, Cc
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta, p.r_A = 3.0, 0.7, 0.04, 0.7, 0.5
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta, p.b_A = 1.0, 0.1, 0.04, -0.7, 0.5
    p.sigma, p.delta = 0.1, 0.1
    p.Fg[0], p.Fg[1] = 6.25e-6, 0.0
    return p


def case_mrt_csf():
    """test/mrt_rayleigh_taylor.cpp: only runs at 1024 x 256 (E_rep is hard-wired, :180).  Stored: a strided view
    (every 8th row, 4th column) of rho, u, phase and the interfacial tension (saved as gradx / grady, :485-486)."""
    d = workdir("mrt_csf")
    toml = os.path.join(d, "csf.toml")
    NS, R, Cc = 30, 1024, 256
    open(toml, "w").write(MRTCG_TOML.format(
        GENERAL='\n[general]\nsigma = 0.1\ngravity_magnitude = 6.25e-6\nname = "g"\n', R=R, C=Cc, T=NS))
    if not os.path.exists(os.path.join(d, "g-mrtcg-rayleigh-taylor-grady.pt")):
        r = run_ref_driver("mrt_rayleigh_taylor", [toml], d, timeout=7200)
        assert r.returncode == 0, r.stderr[-2000:]
    pre = os.path.join(d, "g-mrtcg-rayleigh-taylor-")
    ref = {k: load_pt(pre + k + ".pt") for k in ("rhos", "uxs", "uys", "phases", "gradx", "grady")}
    p = csf_params(R, Cc)
    st = ORC.csf_init(p)
    worst = 0.0
    keep = [0, 1, 2, 3, 5, 10, 20, NS - 1]
    sr, sc = 8, 4
    out = {k: [] for k in ref}
    tols = []
    for t in range(NS):
        # snapshot t: rho, u at the START of iteration t; phase and interf_tension of iteration t-1
        e = max(float(np.abs(st["rho"][..., 0] - ref["rhos"][..., t]).max()), float(np.abs(st["u"][..., 0] - ref["uxs"][..., t]).max()),
                float(np.abs(st["u"][..., 1] - ref["uys"][..., t]).max()), float(np.abs(st["phase"][..., 0] - ref["phases"][..., t]).max()),
                float(np.abs(st["Fs"][..., 0] - ref["gradx"][..., t]).max()), float(np.abs(st["Fs"][..., 1] - ref["grady"][..., t]).max()))
        worst = max(worst, e)
        if t in keep:
            print(f"  snapshot {t}: oracle vs driver max abs err {e:.2e}")
            tols.append(max(1e-12, 10.0 * worst))
            for k in ref:
                out[k].append(ref[k][::sr, ::sc, t])
        ORC.csf_step(p, st)
    # The normal field n = -grad / (1e-20 + |grad|) is O(1) rounding noise wherever the phase gradient is (:508), and
    # the curvature differentiates it: from the third step on the driver and the oracle differ by ~1e-7 in the
    # interfacial tension at the rim of the interface.  The first three snapshots pin the arithmetic at 1e-12; the
    # later ones carry the tolerance the oracle itself meets (x10).
    assert tols[2] <= 1e-12 and worst < 1e-5, (tols, worst)
    save("mrt_csf_1024x256", steps=np.array(keep), stride=np.array([sr, sc]), tol=np.array(tols),
         **{k: np.stack(v) for k, v in out.items()}, toml=open(toml).read())


# ------------------------------------------------------------------ driver 18
def case_mrtcg_droplet():
    d = workdir("mrtcg_sd")
    toml = os.path.join(d, "sd.toml")
    NS, R, Cc = 60, 72, 72
    open(toml, "w").write(MRTCG_TOML.format(GENERAL="", R=R, C=Cc, T=NS))
    r = run_ref_driver("mrtcg_static_droplet", [toml], d)
    assert r.returncode == 0, r.stderr
    p = mrtcg_params(R, Cc, 0.1, (0.0, -6.25e-6), 0)
    worst, init, ref = compare_mrtcg("mrtcg-static-droplet-", d, p, "droplet", NS)
    print(f"  oracle vs reference driver over {NS} snapshots ({R}x{Cc}): worst abs err {worst:.3e}")
    # Not rounding-level: at the droplet centre grad(phase) is pure summation noise (~1e-16) and
    # the recolouring term uses its DIRECTION grad/(1e-20+|grad|), so the minority density there
    # (rho_blue ~ 3e-11) moves by O(rho_blue) with the conv2d summation order.  Bounded by ~1e-11.
    assert worst < 1e-10
    keep = [0, 1, 2, 5, 10, 30, NS - 1]
    save("mrtcg_droplet_72x72", steps=np.array(keep), **{k: np.stack([v[..., t] for t in keep]) for k, v in ref.items()},
         toml=open(toml).read())


# ------------------------------------------------------------------ driver 17
def case_rk():
    d = workdir("rk")
    if not os.path.exists(os.path.join(d, "rk-static-droplet-omegas3.pt")):
        r = run_ref_driver("rk_static_droplet_test", [], d, timeout=7200)
        assert r.returncode == 0, r.stderr
    p = RkParams()
    p.L, p.radius = 101, 25.0
    p.r_rho0, p.r_alpha, p.r_A, p.r_nu = 1.2, 1.0 / 3.0, 1e-4, 0.16
    p.b_rho0, p.b_alpha, p.b_A, p.b_nu = 1.0, 0.2, 1e-4, 0.14
    p.delta = 0.98
    NS = 300
    names = dict(r_fs="r-fs", b_fs="b-fs", ux="ux", uy="uy", rho="rho", rhon="rhon", gradx="gradx", grady="grady",
                 rparams="rparams")
    # the diagnostic fields of every iteration (:546-600): functions of the state at the top of the iteration
    diag = dict(nx="nx", ny="ny", ks="ks", norms="norms", fx="fx", fy="fy", kappas="kappas", omegas1="omegas1",
                omegas2="omegas2", omegas3="omegas3")
    ref = {}
    for k, n in {**names, **diag}.items():
        a = load_pt(os.path.join(d, f"rk-static-droplet-{n}.pt"))
        ref[k] = np.ascontiguousarray(a[..., :NS]).copy()
        del a
    # the driver seeds u with unseeded 1e-15 gaussian noise (:486-487); the oracle starts from u = 0,
    # so agreement is to ~1e-15 absolute, not bit-exact
    st = ORC.rk_init(p)
    worst = worst_diag = 0.0
    for t in range(NS):
        dg = ORC.rk_diagnostics(p, st)
        got = dict(nx=dg["n"][..., 0], ny=dg["n"][..., 1], ks=dg["K"], norms=dg["norm"], fx=dg["Fs"][..., 0], fy=dg["Fs"][..., 1],
                   kappas=dg["kappa"], omegas1=dg["omega1"], omegas2=dg["omega2"], omegas3=dg["omega3"])
        worst_diag = max([worst_diag] + [float(np.abs(got[k] - ref[k][..., t]).max()) for k in got])
        ORC.rk_step(p, st)
        worst = max(worst, float(np.abs(st["r_adv"] - ref["r_fs"][..., t]).max()), float(np.abs(st["b_adv"] - ref["b_fs"][..., t]).max()),
                    float(np.abs(st["u"][..., 0] - ref["ux"][..., t]).max()), float(np.abs(st["u"][..., 1] - ref["uy"][..., t]).max()),
                    float(np.abs(st["rho"] - ref["rho"][..., t]).max()), float(np.abs(st["phase"] - ref["rhon"][..., t]).max()),
                    float(np.abs(st["grad"][..., 0] - ref["gradx"][..., t]).max()), float(np.abs(st["grad"][..., 1] - ref["grady"][..., t]).max()),
                    float(np.abs(st["relax"] - ref["rparams"][..., t]).max()))
    print(f"  oracle vs reference driver over {NS} steps (101x101): worst abs err {worst:.3e}, diagnostic fields {worst_diag:.3e}")
    assert worst < 1e-12 and worst_diag < 1e-12
    keep = [0, 1, 10, NS - 1]
    # omegas3 = omegas1 + omegas2 (eval_omega3 :232-236) was checked above and is not stored
    save("rk_droplet_101", steps=np.array(keep), **{k: np.stack([v[..., t] for t in keep]) for k, v in ref.items() if k != "omegas3"})


# ------------------------------------------------------------------ driver 26 (ulbm::d2q9::kbc)
def case_kbc_double_shear():
    """test/ulbm_double_shear_flow.cpp as shipped: 128 x 128, 10^4 steps, snapshots of m1, m0 every 10 steps.
    Entry ts of the saved stacks holds the fields at the START of iteration 10 ts.  The flow (Re ~ 15000 shear
    layers) amplifies rounding differences, so every stored snapshot carries the tolerance the C oracle itself
    meets against the driver (x10 margin): that is the tolerance the CUDA path is held to."""
    d = workdir("kbc_dsf")
    if not os.path.exists(os.path.join(d, "ulbm-double-shear-flow-rho.pt")):
        r = run_ref_driver("ulbm_double_shear_flow", [], d, timeout=7200)
        assert r.returncode == 0, r.stderr
    ux = load_pt(os.path.join(d, "ulbm-double-shear-flow-ux.pt"))
    uy = load_pt(os.path.join(d, "ulbm-double-shear-flow-uy.pt"))
    rho = load_pt(os.path.join(d, "ulbm-double-shear-flow-rho.pt"))
    R, Cc, Ts = ux.shape
    nu = 1.70766666e-4
    s2 = 1.0 / (0.5 + 3.0 * nu)
    r_ = np.arange(R)[:, None] + 0.0 * np.arange(Cc)[None, :]
    c_ = np.arange(Cc)[None, :] + 0.0 * np.arange(R)[:, None]
    u = np.zeros((R, Cc, 2))
    u[..., 0] = 0.02 * np.tanh(80.0 * (0.25 * R - np.abs(c_ - 0.5 * R)))
    u[..., 1] = 0.02 * 0.05 * np.sin(6.2832 * (r_ + 0.25 * R) / R)
    m0 = np.ones((R, Cc))
    assert np.array_equal(ux[..., 0], u[..., 0]) and np.array_equal(uy[..., 0], u[..., 1])
    f = ORC.kbc_equilibrium(m0, u, fresh_object=True)
    keep = [10, 100, 1000, 9990]
    st = 2  # every second row and column is stored (fixture size)
    out = dict(steps=[], ux=[], uy=[], rho=[], tol=[])
    for t in range(1, keep[-1] + 1):
        ORC.kbc_step(f, m0, u, s2)
        if t in keep:
            ts = t // 10
            e = max(np.abs(u[..., 0] - ux[..., ts]).max(), np.abs(u[..., 1] - uy[..., ts]).max(), np.abs(m0 - rho[..., ts]).max())
            print(f"  step {t}: oracle vs driver max abs err {e:.2e}")
            out["steps"].append(t); out["ux"].append(ux[::st, ::st, ts]); out["uy"].append(uy[::st, ::st, ts]); out["rho"].append(rho[::st, ::st, ts])
            out["tol"].append(max(1e-12, 10.0 * e))
    assert out["tol"][0] <= 1e-12 and out["tol"][1] <= 1e-11, out["tol"]
    save("kbc_double_shear_128", steps=np.array(out["steps"]), ux=np.array(out["ux"]), uy=np.array(out["uy"]),
         rho=np.array(out["rho"]), tol=np.array(out["tol"]), s2=s2, stride=st)


CASES = dict(loop_blocks=case_loop_blocks, mrt_csf=case_mrt_csf, kbc_double_shear=case_kbc_double_shear, poiseuille=case_poiseuille, specular=case_specular, gravity=case_gravity, decompose=case_decompose,
             free_stream=case_free_stream, cylinder=case_cylinder, sedimentation=case_sedimentation,
             mrtcg_rt=case_mrtcg_rt, mrtcg_droplet=case_mrtcg_droplet, rk=case_rk)

if __name__ == "__main__":
    todo = sys.argv[1:] or list(CASES)
    for c in todo:
        print(f"[{c}]")
        CASES[c]()
