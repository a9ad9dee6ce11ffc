"""CPU: the product's collision arithmetic compiled for the host (tests/host_kernels/host_kernels.cu: the very
__host__ __device__ functions of csrc/lbm_device.cuh that the kernels inline) against the oracle.  Lets a change of the
arithmetic be checked here, where there is no GPU; the GPU suite then only has to confirm the fused-multiply-add build."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases
from oracle_lib import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_kernels", "host_kernels.cu")
OUT = os.path.join(HERE, "host_kernels", "_build", "libhost_kernels.so")
CSRC = os.path.join(os.path.dirname(HERE), "lattice-boltzmann-method_b200", "csrc")
DEPS = [SRC, os.path.join(CSRC, "lbm_device.cuh"), os.path.join(CSRC, "lbm_two_phase.cuh")]
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

CXI = (0, 1, 0, -1, 0, 1, -1, -1, 1)
CYI = (0, 0, 1, 0, -1, 1, 1, -1, -1)


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    dp = C.POINTER(C.c_double)
    lib.host_kbc_collide.argtypes = [dp, dp, dp, C.c_int, C.c_long, C.c_double]
    return lib


def ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def advect(f):
    """solver::advect / kbc::advect: periodic streaming"""
    out = np.empty_like(f)
    for q in range(9):
        out[..., q] = np.roll(f[..., q], (CXI[q], CYI[q]), axis=(0, 1))
    return out


@pytest.mark.parametrize("given", [False, True])
def test_kbc_collision_on_the_host_matches_the_oracle(host, given):
    """double shear layer (test/ulbm_double_shear_flow.cpp) with noise on the populations, 12 steps: collision from the
    product's source, streaming in numpy, against orc_kbc_step; `given`: the caller's m0, u differ from the populations'"""
    orc = Oracle()
    R = 48
    s2 = 1.0 / (0.5 + 3.0 * 1.70766666e-4)
    m0, u = cases.double_shear_fields(R, R)
    f = orc.kbc_equilibrium(m0, u, fresh_object=True)
    f *= 1.0 + 1e-3 * np.random.default_rng(3).standard_normal(f.shape)
    if given:
        m0 = m0 * (1.0 + 1e-4 * np.random.default_rng(4).standard_normal(m0.shape))
        u = u + 1e-5 * np.random.default_rng(5).standard_normal(u.shape)
    else:
        m0 = f.sum(-1)
        u = np.stack([(f * np.array(CXI)).sum(-1), (f * np.array(CYI)).sum(-1)], -1) / m0[..., None]
    ref_f, ref_m0, ref_u = f.copy(), np.ascontiguousarray(m0), np.ascontiguousarray(u)
    mine = f.copy()
    mine_m0, mine_u = ref_m0.copy(), ref_u.copy()
    for step in range(12):
        orc.kbc_step(ref_f, ref_m0, ref_u, s2)
        c = np.ascontiguousarray(mine)
        host.host_kbc_collide(ptr(c), ptr(mine_m0), ptr(mine_u), 1 if (given and step == 0) else 0, c.size // 9, s2)
        mine = advect(c)
        assert cases.relerr(mine, ref_f) < 1e-13, step


# ------------------------------------------------------------------------------------------------
# two-phase models: tp_moments / tp_collide / tp_fill_params of csrc/lbm_two_phase.cuh on the host
# ------------------------------------------------------------------------------------------------
TP_MRTCG, TP_RK, TP_CSF = 0, 1, 2   # enum TpModel of csrc/lbm_two_phase.cuh


def tp_api(lib):
    import lbm_b200 as L
    dp = C.POINTER(C.c_double)
    cp = C.POINTER(L.Config)
    lib.host_tp_moments.argtypes = [C.c_int, cp, dp, dp, dp, C.c_long, dp, dp, dp, dp]
    lib.host_tp_phase.argtypes = [C.c_int, cp, dp, dp, C.c_long, dp]
    lib.host_tp_collide.argtypes = [C.c_int, cp, dp, dp, dp, dp, dp, dp, dp, dp, C.c_long]
    lib.host_tp_constants.argtypes = [C.c_int, cp, dp]
    return L


def mrtcg_bc(adv, col):
    """test/mrtcg_rayleigh_taylor.cpp:495-533: periodic columns on the inner rows, bounce-back on the first / last row"""
    X, Y = adv.shape[:2]
    inner = slice(1, X - 1)
    for q in (2, 5, 6):
        adv[inner, 0, q] = col[inner, Y - 1, q]
    for q in (4, 8, 7):
        adv[inner, Y - 1, q] = col[inner, 0, q]
    for q, qs in ((3, 1), (7, 5), (6, 8)):
        adv[X - 1, :, q] = col[X - 1, :, qs]
    for q, qs in ((1, 3), (5, 7), (8, 6)):
        adv[0, :, q] = col[0, :, qs]


def rk_bc(adv, col):
    """test/rk_static_droplet_test.cpp:204-211: whole nodes copied across both pairs of edges"""
    X, Y = adv.shape[:2]
    inner = slice(1, X - 1)
    adv[inner, 0, :] = col[inner, Y - 1, :]
    adv[inner, Y - 1, :] = col[inner, 0, :]
    adv[0, :, :] = col[X - 1, :, :]
    adv[X - 1, :, :] = col[0, :, :]


def tp_emulated_step(host, orc, L, model, cfg, fr, fb, rr, rb, u, Fs=None):
    """one step of a two-phase model: the product's moments-to-collision arithmetic on the host, the differences through
    the oracle's operators, streaming and the drivers' boundary rules in numpy.  Returns the new populations and moments."""
    shape = rr.shape
    N = rr.size
    k = np.zeros(10)
    host.host_tp_constants(model, C.byref(cfg), ptr(k))
    ph = np.zeros(shape)
    host.host_tp_phase(model, C.byref(cfg), ptr(rr), ptr(rb), N, ptr(ph))
    st4 = np.zeros(shape + (4,))
    if model == TP_RK:
        st4[..., 0], st4[..., 1] = orc.diff3(ph)
    else:
        cq = k[0] * rr + k[1] * rb
        st4[..., 0], st4[..., 1] = orc.diff5(ph)
        st4[..., 2] = orc.diff5(np.ascontiguousarray(cq * u[..., 0]))[0]
        st4[..., 3] = orc.diff5(np.ascontiguousarray(cq * u[..., 1]))[1]
    cr, cb = np.ascontiguousarray(fr).copy(), np.ascontiguousarray(fb).copy()
    host.host_tp_collide(model, C.byref(cfg), ptr(cr), ptr(cb), ptr(rr), ptr(rb), ptr(np.ascontiguousarray(u)), ptr(ph), ptr(st4),
                         ptr(Fs) if Fs is not None else None, N)
    nr, nb = advect(cr), advect(cb)
    (rk_bc if model == TP_RK else mrtcg_bc)(nr, cr)
    (rk_bc if model == TP_RK else mrtcg_bc)(nb, cb)
    nr, nb = np.ascontiguousarray(nr), np.ascontiguousarray(nb)
    rr2, rb2, u2, ph2 = np.zeros(shape), np.zeros(shape), np.zeros(shape + (2,)), np.zeros(shape)
    host.host_tp_moments(model, C.byref(cfg), ptr(nr), ptr(nb), None, N, ptr(rr2), ptr(rb2), ptr(u2), ptr(ph2))
    return nr, nb, rr2, rb2, u2


@pytest.mark.parametrize("kind", ["rt", "droplet"])
def test_mrtcg_collision_on_the_host_matches_the_oracle(host, kind):
    """MRT colour gradient (test/mrtcg_rayleigh_taylor.cpp, mrtcg_static_droplet.cpp), 10 steps from the drivers' own
    initial states: populations and velocity at 1e-12 (the droplet's recolouring divides by a noise-level gradient at the
    centre of the drop: 1e-10 there, as for the oracle itself against the reference, tests/golden/make_golden.py)"""
    from oracle_lib import MrtcgParams
    L = tp_api(host)
    orc = Oracle()
    R, Cc = (48, 40) if kind == "rt" else (56, 56)
    Fg, add_force = ((6.25e-6, 0.0), 1) if kind == "rt" else ((0.0, -6.25e-6), 0)
    p = MrtcgParams()
    p.R, p.C = R, Cc
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = 3.0, 0.7, 0.04, 0.7
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = 1.0, 0.1, 0.04, -0.7
    p.sigma, p.delta = 0.1, 0.1
    p.Fg[0], p.Fg[1] = Fg
    p.add_force = add_force
    st = orc.mrtcg_init(p, kind)
    cfg = L.default_config(model=L.MODEL_MRTCG, X=R, Y=Cc, red=cases.RED, blue=cases.BLUE, sigma=0.1, delta=0.1, Fg=Fg,
                           add_force=add_force)
    fr, fb = st["r_adv"].copy(), st["b_adv"].copy()
    rr, rb, u = st["r_rho"][..., 0].copy(), st["b_rho"][..., 0].copy(), st["u"].copy()
    tol = 1e-12 if kind == "rt" else 1e-10
    for step in range(10):
        fr, fb, rr, rb, u = tp_emulated_step(host, orc, L, TP_MRTCG, cfg, fr, fb, rr, rb, u)
        orc.mrtcg_step(p, st)
        assert cases.relerr(fr, st["r_adv"]) < tol and cases.relerr(fb, st["b_adv"]) < tol, step
        assert np.abs(u - st["u"]).max() < tol and np.abs(rr - st["r_rho"][..., 0]).max() < tol, step


def test_rk_collision_on_the_host_matches_the_oracle(host):
    """Rothman-Keller static droplet (test/rk_static_droplet_test.cpp), 10 steps"""
    from oracle_lib import RkParams
    L = tp_api(host)
    orc = Oracle()
    Ln = 48
    p = RkParams()
    p.L, p.radius = Ln, Ln / 4.0
    p.r_rho0, p.r_alpha, p.r_A, p.r_nu = 1.2, 1.0 / 3.0, 1e-4, 0.16
    p.b_rho0, p.b_alpha, p.b_A, p.b_nu = 1.0, 0.2, 1e-4, 0.14
    p.delta = 0.98
    st = orc.rk_init(p)
    cfg = L.default_config(model=L.MODEL_RK, X=Ln, Y=Ln, red=cases.RK_RED, blue=cases.RK_BLUE, delta=0.98)
    fr, fb = st["r_adv"].copy(), st["b_adv"].copy()
    rr, rb, u = st["r_rho"].copy(), st["b_rho"].copy(), st["u"].copy()
    for step in range(10):
        fr, fb, rr, rb, u = tp_emulated_step(host, orc, L, TP_RK, cfg, fr, fb, rr, rb, u)
        orc.rk_step(p, st)
        assert cases.relerr(fr, st["r_adv"]) < 1e-12 and cases.relerr(fb, st["b_adv"]) < 1e-12, step
        assert np.abs(u - st["u"]).max() < 1e-12 and np.abs(rr - st["r_rho"]).max() < 1e-12, step


def test_csf_collision_on_the_host_matches_the_oracle(host):
    """MRT colour gradient with continuum surface force (test/mrt_rayleigh_taylor.cpp), 6 steps from the driver's sharp
    interface: normals, curvature and interfacial tension assembled here from the oracle's difference operator (the same
    routine the oracle's own step uses, so the rounding-residue normals of the bulk agree), collision and moments from the
    product's source"""
    from oracle_lib import CsfParams
    L = tp_api(host)
    orc = Oracle()
    R, Cc = 64, 40
    p = CsfParams()
    p.R, p.C = R, Cc
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta, p.r_A = 3.0, 0.7, 0.04, 0.7, 0.5
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta, p.b_A = 1.0, 0.1, 0.04, -0.7, 0.5
    p.sigma, p.delta = 0.1, 0.1
    p.Fg[0], p.Fg[1] = 6.25e-6, 0.0
    st = orc.csf_init(p)
    cfg = L.default_config(model=L.MODEL_MRT_CSF, X=R, Y=Cc, red=cases.RED, blue=cases.BLUE, sigma=0.1, delta=0.1,
                           Fg=(6.25e-6, 0.0), add_force=1)
    k = np.zeros(10)
    host.host_tp_constants(TP_CSF, C.byref(cfg), ptr(k))
    fr, fb = st["r_adv"].copy(), st["b_adv"].copy()
    rr, rb, u = st["r_rho"][..., 0].copy(), st["b_rho"][..., 0].copy(), st["u"].copy()
    N = rr.size
    for step in range(6):
        ph = np.zeros((R, Cc))
        host.host_tp_phase(TP_CSF, C.byref(cfg), ptr(rr), ptr(rb), N, ptr(ph))
        gx, gy = orc.diff5(ph)
        gn = np.sqrt(gx * gx + gy * gy)
        nx, ny = np.ascontiguousarray(-gx / (1e-20 + gn)), np.ascontiguousarray(-gy / (1e-20 + gn))
        dx_nx, dy_nx = orc.diff5(nx)
        dx_ny, dy_ny = orc.diff5(ny)
        K = nx * ny * (dy_nx + dx_ny) - (nx * nx) * dy_ny - (ny * ny) * dx_nx      # eval_local_curvature (:355-364)
        Fs = np.ascontiguousarray(np.stack([(-0.5 * 0.1) * K * gx, (-0.5 * 0.1) * K * gy], -1))   # (:510)
        cq = k[0] * rr + k[1] * rb
        st4 = np.zeros((R, Cc, 4))
        st4[..., 0], st4[..., 1] = gx, gy
        st4[..., 2] = orc.diff5(np.ascontiguousarray(cq * u[..., 0]))[0]
        st4[..., 3] = orc.diff5(np.ascontiguousarray(cq * u[..., 1]))[1]
        cr, cb = np.ascontiguousarray(fr).copy(), np.ascontiguousarray(fb).copy()
        host.host_tp_collide(TP_CSF, C.byref(cfg), ptr(cr), ptr(cb), ptr(rr), ptr(rb), ptr(np.ascontiguousarray(u)), ptr(ph),
                             ptr(st4), ptr(Fs), N)
        fr, fb = advect(cr), advect(cb)
        mrtcg_bc(fr, cr)
        mrtcg_bc(fb, cb)
        fr, fb = np.ascontiguousarray(fr), np.ascontiguousarray(fb)
        rr, rb, u = np.zeros((R, Cc)), np.zeros((R, Cc)), np.zeros((R, Cc, 2))
        host.host_tp_moments(TP_CSF, C.byref(cfg), ptr(fr), ptr(fb), ptr(Fs), N, ptr(rr), ptr(rb), ptr(u), ptr(ph))
        orc.csf_step(p, st)
        tol = 1e-12 if step < 2 else 1e-9   # later steps: the algorithm's own conditioning
        assert cases.relerr(fr, st["r_adv"]) < tol and cases.relerr(fb, st["b_adv"]) < tol, step
        assert np.abs(u - st["u"]).max() < tol and np.abs(Fs - st["Fs"]).max() < tol, step


# ------------------------------------------------------------------------------------------------
# single-phase family: bgk_collide<EQ, FORCE> on the host
# ------------------------------------------------------------------------------------------------
W9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def bgk_api(lib):
    dp = C.POINTER(C.c_double)
    lib.host_bgk_collide.argtypes = [C.c_int, C.c_int, dp, C.c_long, C.c_double, C.c_double, C.c_double, dp, C.c_double,
                                     C.c_double, dp, dp]


def near_equilibrium(orc, X, Y, seed):
    rng = np.random.default_rng(seed)
    rho = 1.0 + 0.02 * rng.standard_normal((X, Y, 1))
    u = 0.05 * rng.standard_normal((X, Y, 2))
    return orc.equilibrium(u, rho) * (1.0 + 1e-2 * rng.standard_normal((X, Y, 9)))


@pytest.mark.parametrize("eq", [0, 1])
def test_bgk_collision_on_the_host_matches_the_operators(host, eq):
    """solver::calc_rho / calc_u / calc_incomp_u / equilibrium / incomp_equilibrium / collision fused in bgk_collide"""
    bgk_api(host)
    orc = Oracle()
    f = near_equilibrium(orc, 24, 20, 1)
    omega = 1.37
    rho = orc.calc_rho(f)
    u = orc.calc_u(f, rho) if eq == 0 else orc.calc_incomp_u(f)
    want = orc.collision(f, orc.equilibrium(u, rho) if eq == 0 else orc.incomp_equilibrium(u, rho), omega)
    got = f.copy()
    r, uu = np.zeros(f.shape[:2]), np.zeros(f.shape[:2] + (2,))
    host.host_bgk_collide(eq, 0, ptr(got), got.size // 9, omega, 0.0, 0.0, None, 0.0, 0.0, ptr(r), ptr(uu))
    assert cases.relerr(got, want) < 1e-15
    assert np.abs(r - rho[..., 0]).max() < 1e-15 and np.abs(uu - u).max() < 1e-15


@pytest.mark.parametrize("force,ics2,ics4", [(2, 1.0 / 3.0, 1.0 / 9.0), (3, 3.0, 9.0)])
def test_bgk_force_field_source_term_on_the_host(host, force, ics2, ics4):
    """f + (-omega (f - feq)) + (1 - omega/2) ((ics2 + ics4 u.c)(F.c) - ics2 u.F) w: cylinder_test.cpp:112-127 with 1/3, 1/9
    (compile-time in the kernels) and decompose_domain_loop.cpp:66-69,151-158 with 3, 9 (lbm_set_force_region)"""
    bgk_api(host)
    orc = Oracle()
    f = near_equilibrium(orc, 20, 18, 2)
    omega = 1.21
    F = np.ascontiguousarray(1e-3 * np.random.default_rng(9).standard_normal(f.shape[:2] + (2,)))
    rho = orc.calc_rho(f)
    u = orc.calc_u(f, rho)
    feq = orc.equilibrium(u, rho)
    CX = np.array(CXI, dtype=float); CY = np.array(CYI, dtype=float)
    cu = u[..., 0:1] * CX + u[..., 1:2] * CY
    cF = F[..., 0:1] * CX + F[..., 1:2] * CY
    uF = u[..., 0:1] * F[..., 0:1] + u[..., 1:2] * F[..., 1:2]
    want = f + (-omega * (f - feq)) + ((1.0 - 0.5 * omega) * ((ics2 + ics4 * cu) * cF - ics2 * uF)) * W9
    got = f.copy()
    r, uu = np.zeros(f.shape[:2]), np.zeros(f.shape[:2] + (2,))
    host.host_bgk_collide(0, force, ptr(got), got.size // 9, omega, 0.0, 0.0, ptr(F), ics2, ics4, ptr(r), ptr(uu))
    assert cases.relerr(got, want) < 1e-15


def test_bgk_uniform_force_on_the_host_matches_the_gravity_driver(host):
    """test/gravity_test.cpp: one collision of the oracle's gravity step (u += Fg before the equilibrium, source term with
    1/3 and 1/9), isolated by undoing the streaming on a periodic box away from the driver's boundary rows / columns"""
    bgk_api(host)
    orc = Oracle()
    X = Y = 16
    f = near_equilibrium(orc, X, Y, 3)
    omega, Fg = 1.1, (-3e-4, 1e-4)
    got = f.copy()
    r, uu = np.zeros((X, Y)), np.zeros((X, Y, 2))
    host.host_bgk_collide(1, 1, ptr(got), X * Y, omega, Fg[0], Fg[1], None, 0.0, 0.0, ptr(r), ptr(uu))
    ref = f.copy()
    u = np.zeros((X, Y, 2)); rho = np.ones((X, Y, 1))
    orc.gravity_step(ref, u, rho, omega, 1.0, 1.0, np.array(Fg))
    mine = advect(got)
    inner = (slice(3, X - 3), slice(3, Y - 3))     # nodes whose nine sources no boundary rule of the driver touches
    assert cases.relerr(mine[inner], ref[inner]) < 1e-14
