"""Parity at the grid sizes bench.py times (8192^2, 16384^2, 4096^2) — the index paths no small case reaches: plane
offsets beyond 2^31 doubles, 64 .. 133 column strips, 64 .. 128 row bands, hundreds of "early" rows.

The CPU oracle cannot run those grids in test time, so every test uses the one size-independent property the models offer:
the update of a node depends only on its neighbourhood.  The grid is filled with a 128 x 128 tile repeated in both
directions.  Until the influence of the grid's edges (walls, inlet rows, replicate padding, an immersed body) has travelled
one tile inwards — 1 cell per step for the single-phase models, 3 for the MRT colour gradient (streaming + the 5x5
differences), 2 for Rothman-Keller, 5 for the CSF variant — every tile that is at least one tile away from those features
(a) equals every other such tile BIT FOR BIT, whichever strip, band or slab computed it, and
(b) equals the middle tile of the ORACLE's run on a 3 x 3-tile grid (384 x 384) within the stated tolerance: 1e-12 relative
    on populations / 1e-9 on rho, u, phase (north_star's bars).
Plus: an 8192^2 domain cut into three linked slabs with a body across a cut equals the monolithic run bit for bit."""
import ctypes
import os

import numpy as np
import pytest

import cases
import lbm_b200 as L
from oracle_lib import CsfParams, MrtcgParams, Oracle, RkParams

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

T = 128  # tile edge
EMU = os.environ.get("LBM_EMU") == "1"  # tests/cpu_emu: the same checks on 4 x 4 tiles (the logic, not the index range)


def full(n):
    return 4 * T if EMU else n


W9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1.0])
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1.0])


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def tile_fields(seed=3):
    """smooth T-periodic rho, u (single-phase) on one tile"""
    x = 2 * np.pi * np.arange(T)[:, None] / T
    y = 2 * np.pi * np.arange(T)[None, :] / T
    rho = 1.0 + 0.02 * np.sin(x + 0.3) * np.cos(2 * y) + 0.01 * np.cos(3 * x - y)
    u = np.zeros((T, T, 2))
    u[..., 0] = 0.03 + 0.02 * np.sin(2 * x) * np.sin(y + 0.7)
    u[..., 1] = 0.015 * np.cos(x - 2 * y)
    return rho[..., None], u


def tile_two_phase(r0, b0):
    """a T-periodic diffuse blob of the red fluid in the blue one"""
    x = np.arange(T)[:, None] - T / 2 + 0.5
    y = np.arange(T)[None, :] - T / 2 + 0.5
    s = np.sqrt(x * x + (1.3 * y) ** 2)
    w = 0.5 * (1.0 - np.tanh((s - 30.0) / 5.0))
    return r0 * w, b0 * (1.0 - w)


def tiled(a, nx, ny):
    return np.tile(a, (nx, ny) + (1,) * (a.ndim - 2))


def assert_tiles_identical(a, margin, what):
    """a: {X,Y,...}; every tile at least `margin` tiles away from the grid's edges equals tile (margin, margin) bit for bit"""
    nx, ny = a.shape[0] // T, a.shape[1] // T
    ref = a[margin * T:(margin + 1) * T, margin * T:(margin + 1) * T]
    for i in range(margin, nx - margin):  # row of tiles at a time: no grid-sized temporaries
        row = a[i * T:(i + 1) * T, margin * T:(ny - margin) * T]
        blocks = row.reshape((T, ny - 2 * margin, T) + a.shape[2:])
        same = (blocks == ref[:, None]).all()
        if not same:
            bad = np.argwhere(~(blocks == ref[:, None]).reshape(T, ny - 2 * margin, T, -1).all(axis=(0, 2, 3)))
            raise AssertionError(f"{what}: tile row {i}, tile columns {bad[:8].ravel() + margin} differ from tile ({margin},{margin})")
    return ref


def test_bgk_cylinder_kernel_at_8192(orc):
    """the headline instantiation k_bgk_interior<PULL,COMP,IBM> at the benched size: free-stream rules, a small immersed body
    in the first tile (its ROI rows become early rows), 50 steps"""
    X = Y = full(8192)
    omega, u_lb, steps = 1.0 / 0.55, 0.03, 50
    rho_t, u_t = tile_fields()
    f_t = orc.equilibrium(u_t, rho_t)
    th = 2 * np.pi * np.arange(40) / 40
    xs, ys = 40.3 + 9.0 * np.cos(th), 50.2 + 9.0 * np.sin(th)
    d = cases.cylinder(X, Y, omega, u_lb, xs, ys)
    d.set_f(tiled(f_t, X // T, Y // T))
    d.step(steps)
    got = d.get_f()
    d.close()
    ref = assert_tiles_identical(got, 1, "populations")
    # oracle: 3 x 3 tiles, the same rules, the same body (in the corner tile, out of reach of the middle one)
    f = tiled(f_t, 3, 3)
    u = np.zeros((3 * T, 3 * T, 2)); rho = np.ones((3 * T, 3 * T, 1))
    ib = orc.ibm_create(xs, ys)
    for _ in range(steps):
        orc.cylinder_step(f, u, rho, omega, u_lb, ib)
    orc.ibm_destroy(ib)
    assert cases.relerr(ref, f[T:2 * T, T:2 * T]) < 1e-12


def test_kbc_at_8192(orc):
    """fully periodic double-shear-like field: EVERY tile, the grid's edge tiles included, equals the oracle's 128 x 128 run"""
    X = Y = full(8192)
    s2, steps = 1.0 / (0.5 + 3.0 * 1.70766666e-4), 40
    rho_t, u_t = tile_fields()
    m0 = np.ascontiguousarray(rho_t[..., 0])
    f_t = orc.kbc_equilibrium(m0, u_t)
    d = cases.kbc(X, Y, s2)
    d.set_f(tiled(f_t, X // T, Y // T))
    d.set_moments(tiled(rho_t, X // T, Y // T), tiled(u_t, X // T, Y // T))
    d.step(steps)
    got = d.get_f()
    d.close()
    ref = assert_tiles_identical(got, 0, "populations")
    f, m, u = f_t.copy(), m0.copy(), u_t.copy()
    for _ in range(steps):
        orc.kbc_step(f, m, u, s2)
    assert cases.relerr(ref, f) < 1e-12


def two_phase_oracle_params(cls, n):
    p = cls()
    if cls is RkParams:
        p.L, p.radius = n, n / 4.0
        p.r_rho0, p.r_alpha, p.r_A, p.r_nu = 1.2, 1.0 / 3.0, 1e-4, 0.16
        p.b_rho0, p.b_alpha, p.b_A, p.b_nu = 1.0, 0.2, 1e-4, 0.14
        p.delta = 0.98
        return p
    p.R, p.C = n, n
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = 3.0, 0.7, 0.04, 0.7
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = 1.0, 0.1, 0.04, -0.7
    if cls is CsfParams:
        p.r_A = p.b_A = 0.5
    else:
        p.add_force = 1
    p.sigma, p.delta = 0.1, 0.1
    p.Fg[0], p.Fg[1] = 6.25e-6, 0.0
    return p


def test_mrtcg_at_16384(orc):
    """BASELINE configs[2]'s grid: 268 M nodes, plane offsets up to 2.4 G doubles, 133 strips x 128 bands.  rho, u of every
    interior tile after 20 steps (the populations would be 2 x 19 GB on the host)"""
    X = Y = full(16384)
    steps = 20
    rr_t, rb_t = tile_two_phase(3.0, 1.0)
    d = cases.mrtcg(X, Y, (6.25e-6, 0.0), 1)
    d.init_two_phase(tiled(rr_t, X // T, Y // T), tiled(rb_t, X // T, Y // T), np.zeros((X, Y, 2)))
    d.step(steps)
    rho, u = d.get_moments()
    d.close()
    ref_rho = assert_tiles_identical(rho, 1, "rho")
    ref_u = assert_tiles_identical(u, 1, "u")
    del rho, u
    n = 3 * T
    p = two_phase_oracle_params(MrtcgParams, n)
    st = orc.mrtcg_init(p, "rt")
    st["r_rho"][..., 0], st["b_rho"][..., 0] = tiled(rr_t, 3, 3), tiled(rb_t, 3, 3)
    orc.lib.orc_mrtcg_init_state(ctypes.byref(p), st["r_rho"].ctypes.data_as(L.dp), st["b_rho"].ctypes.data_as(L.dp), st["rho"].ctypes.data_as(L.dp),
                                 st["u"].ctypes.data_as(L.dp), st["r_adv"].ctypes.data_as(L.dp), st["b_adv"].ctypes.data_as(L.dp), 0)
    for _ in range(steps):
        orc.mrtcg_step(p, st)
    assert np.abs(ref_rho - st["rho"][T:2 * T, T:2 * T]).max() < 1e-9
    assert np.abs(ref_u - st["u"][T:2 * T, T:2 * T]).max() < 1e-9


def test_mrtcg_populations_at_8192(orc):
    """the same model at 8192^2 with the populations themselves: 1e-12 after one step, 1e-9 after 20"""
    X = Y = full(8192)
    rr_t, rb_t = tile_two_phase(3.0, 1.0)
    d = cases.mrtcg(X, Y, (6.25e-6, 0.0), 1)
    d.init_two_phase(tiled(rr_t, X // T, Y // T), tiled(rb_t, X // T, Y // T), np.zeros((X, Y, 2)))
    n = 3 * T
    p = two_phase_oracle_params(MrtcgParams, n)
    st = orc.mrtcg_init(p, "rt")
    st["r_rho"][..., 0], st["b_rho"][..., 0] = tiled(rr_t, 3, 3), tiled(rb_t, 3, 3)
    orc.lib.orc_mrtcg_init_state(ctypes.byref(p), st["r_rho"].ctypes.data_as(L.dp), st["b_rho"].ctypes.data_as(L.dp), st["rho"].ctypes.data_as(L.dp),
                                 st["u"].ctypes.data_as(L.dp), st["r_adv"].ctypes.data_as(L.dp), st["b_adv"].ctypes.data_as(L.dp), 0)
    done = 0
    for upto, tol in ((1, 1e-12), (20, 1e-9)):
        d.step(upto - done)
        for _ in range(upto - done):
            orc.mrtcg_step(p, st)
        done = upto
        for lat, key in ((0, "r_adv"), (1, "b_adv")):
            got = d.get_f(lat)
            ref = assert_tiles_identical(got, 1, f"lattice {lat} after {upto} steps")
            assert cases.relerr(ref, st[key][T:2 * T, T:2 * T]) < tol, (upto, lat)
            del got
    d.close()


def test_rk_at_4096(orc):
    """BASELINE configs[3]'s grid; populations of both colours after 1 and 30 steps"""
    X = full(4096)
    rr_t, rb_t = tile_two_phase(1.2, 1.0)
    d = cases.rk(X)
    d.init_two_phase(tiled(rr_t, X // T, X // T), tiled(rb_t, X // T, X // T), np.zeros((X, X, 2)))
    n = 3 * T
    p = two_phase_oracle_params(RkParams, n)
    st = orc.rk_init(p)
    small = cases.rk(n)  # the library's own equilibrium import on the oracle's grid: the same initial populations on both sides
    small.init_two_phase(tiled(rr_t, 3, 3), tiled(rb_t, 3, 3), np.zeros((n, n, 2)))
    st["r_adv"][...], st["b_adv"][...] = small.get_f(0), small.get_f(1)
    small.close()
    # the moments the oracle carries from the end of one iteration into the next, of that state (u = 0)
    st["r_rho"][...], st["b_rho"][...] = orc.calc_rho(st["r_adv"])[..., 0], orc.calc_rho(st["b_adv"])[..., 0]
    st["rho"][...] = st["r_rho"] + st["b_rho"]
    st["u"][...] = 0.0
    done = 0
    for upto, tol in ((1, 1e-12), (30, 1e-9)):
        d.step(upto - done)
        for _ in range(upto - done):
            orc.rk_step(p, st)
        done = upto
        for lat, key in ((0, "r_adv"), (1, "b_adv")):
            ref = assert_tiles_identical(d.get_f(lat), 1, f"lattice {lat} after {upto} steps")
            assert cases.relerr(ref, st[key][T:2 * T, T:2 * T]) < tol, (upto, lat)
    d.close()


def test_csf_at_8192(orc):
    """the continuum-surface-force variant (single pass by default) at the benched size; a diffuse interface everywhere in
    the blob's rim, 8 steps (reach 5 cells per step)"""
    X = Y = full(8192)
    steps = 8
    rr_t, rb_t = tile_two_phase(3.0, 1.0)
    d = cases.csf(X, Y)
    d.init_two_phase(tiled(rr_t, X // T, Y // T), tiled(rb_t, X // T, Y // T), np.zeros((X, Y, 2)))
    d.step(steps)
    rho, u = d.get_moments()
    d.close()
    # bulk plateaus turn the interface normal into rounding residue: tiles agree bit for bit all the same —
    # every node sees identical operands in an identical order — but the oracle comparison is on rho, u at 1e-9
    ref_rho = assert_tiles_identical(rho, 1, "rho")
    ref_u = assert_tiles_identical(u, 1, "u")
    del rho, u
    n = 3 * T
    p = two_phase_oracle_params(CsfParams, n)
    st = orc.csf_init(p)
    small = cases.csf(n, n)
    small.init_two_phase(tiled(rr_t, 3, 3), tiled(rb_t, 3, 3), np.zeros((n, n, 2)))
    st["r_adv"][...], st["b_adv"][...] = small.get_f(0), small.get_f(1)
    st["r_rho"][...], st["b_rho"][...] = orc.calc_rho(st["r_adv"]), orc.calc_rho(st["b_adv"])
    st["rho"][...] = st["r_rho"] + st["b_rho"]
    st["u"][...] = 0.0
    small.close()
    for _ in range(steps):
        orc.csf_step(p, st)
    assert np.abs(ref_rho - st["rho"][T:2 * T, T:2 * T]).max() < 1e-9
    assert np.abs(ref_u - st["u"][T:2 * T, T:2 * T]).max() < 1e-9


def test_slabs_equal_monolithic_at_8192():
    """three linked slabs of an 8192^2 cylinder domain, the body's ROI across the first cut: bit-identical to one slab"""
    X = Y = full(8192)
    omega, u_lb, steps = 1.0 / 0.55, 0.03, 12
    rho_t, u_t = tile_fields()
    orc = Oracle()
    f0 = tiled(orc.equilibrium(u_t, rho_t), X // T, Y // T)
    cut = L.decompose_rows(X, 3, 0)[1]
    th = 2 * np.pi * np.arange(600) / 600
    xs, ys = cut + 3.3 + 95.0 * np.cos(th), Y / 2 + 0.2 + 95.0 * np.sin(th)
    mono = cases.cylinder(X, Y, omega, u_lb, xs, ys)
    mono.set_f(f0)
    mono.step(steps)
    want = mono.get_f()
    mono.close()
    slabs = []
    for r in range(3):
        x0, x1 = L.decompose_rows(X, 3, r)
        d = L.Domain(L.default_config(model=L.MODEL_BGK, X=X, Y=Y, x0=x0, x1=x1, omega=omega, equilibrium=L.EQ_COMPRESSIBLE,
                                      force=L.FORCE_IBM))
        d.preset_free_stream(u_lb, 0.0)
        d.ibm_set_markers(xs, ys)
        d.set_f(f0[x0:x1])
        slabs.append(d)
    for r, d in enumerate(slabs):
        d.link(slabs[(r - 1) % 3], slabs[(r + 1) % 3])
    L.step_group(slabs, steps)
    for r, d in enumerate(slabs):
        x0, x1 = L.decompose_rows(X, 3, r)
        assert np.array_equal(d.get_f(), want[x0:x1]), r
        d.close()


def test_sedimentation_ade_at_4096x8192(orc):
    """BASELINE configs[4]'s grid, fluid + advection-diffusion lattice (k_bgk_interior<PULL,COMP,NONE,ADE>; every row of this
    driver is an early row, so the one launch over all rows runs on the side stream): inlet column, extrapolated outlet,
    zero-gradient copies and the rectangle's walls stay within one tile of the grid's edges for 40 steps"""
    X, Y = (full(4096), full(8192)) if not EMU else (4 * T, 4 * T)
    omega, u_lb, w_s, steps = 1.0 / 0.55, 0.02, 3e-3, 40
    walls = (-20, 10, 30)                       # rectangle_sedimentation_test.cpp:73-75: rows from the end, two wall columns
    rho_t, u_t = tile_fields()
    u_t = u_t * 0.5
    u_t[..., 1] += u_lb                         # the driver's stream runs along axis 1
    C_t = 1e-3 * (1.0 + 0.3 * np.sin(2 * np.pi * np.arange(T)[:, None] / T) * np.cos(4 * np.pi * np.arange(T)[None, :] / T))[..., None]
    f_t = orc.equilibrium(u_t, rho_t)
    g_t = orc.equilibrium(u_t + w_s, C_t)       # rectangle_sedimentation_test.cpp:123-131: the sediment settles with u + w_s

    def run_gpu(nx, ny):
        C_w = np.zeros(nx * T); C_w[-10:] = 1e-3
        d = cases.sedimentation(nx * T, ny * T, omega, u_lb, w_s, C_w, walls)
        d.set_f(tiled(f_t, nx, ny), 0)
        d.set_f(tiled(g_t, nx, ny), 1)
        d.step(steps)
        out = d.get_f(0), d.get_f(1)
        d.close()
        return out

    gf, gg = run_gpu(X // T, Y // T)
    ref_f = assert_tiles_identical(gf, 1, "fluid populations")
    ref_g = assert_tiles_identical(gg, 1, "sediment populations")
    del gf, gg
    n = 3 * T
    C_w = np.zeros(n); C_w[-10:] = 1e-3
    f, g = tiled(f_t, 3, 3), tiled(g_t, 3, 3)
    rho = orc.calc_rho(f)
    u = orc.calc_u(f, rho)
    Cc = orc.calc_rho(g)
    for _ in range(steps):
        orc.sedimentation_step(f, g, u, rho, Cc, omega, u_lb, w_s, C_w, *walls)
    assert cases.relerr(ref_f, f[T:2 * T, T:2 * T]) < 1e-12
    assert cases.relerr(ref_g, g[T:2 * T, T:2 * T]) < 1e-12
