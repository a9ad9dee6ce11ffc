"""GPU: the north-star long-run bar — density, velocity and phase fields within 1e-9 of the CPU
oracle after 10^4 time steps — on every ported driver at the reference's own (small) sizes, where
the oracle finishes in seconds.  The oracle itself is pinned against the compiled reference
(tests/test_oracle_golden.py, tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

import cases
import lbm_b200 as L
from oracle_lib import MrtcgParams, Oracle, RkParams

pytestmark = pytest.mark.gpu

STEPS = 10_000
TOL = 1e-9


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def fields_close(d, rho, u, lattice=0):
    rho_g, u_g = d.get_moments(lattice)
    return max(float(np.abs(rho_g.reshape(rho.shape) - rho).max()), float(np.abs(u_g - u).max()))


def test_poiseuille_10k(orc):
    """driver 10 (test/horizontal_poiseuille_test.cpp): rho, u of the state after 10^4 steps"""
    d, (omega, rho_in, rho_out) = cases.poiseuille()
    X = Y = 21
    u = np.zeros((X, Y, 2)); rho = np.ones((X, Y, 1))
    f = orc.incomp_equilibrium(u, rho)
    d.set_f(f)
    for _ in range(STEPS):
        orc.poiseuille_step(f, u, rho, omega, rho_in, rho_out)
    d.step(STEPS)
    # the oracle's u, rho are those of the LAST iteration (moments of the state before its collision);
    # compare the moments of the final populations on both sides instead
    rho_o = orc.calc_rho(f); u_o = orc.calc_incomp_u(f)
    assert fields_close(d, rho_o, u_o) < TOL
    assert cases.relerr(d.get_f(), f) < TOL


def test_specular_channel_10k(orc):
    d, (omega, rho_in, rho_out) = cases.specular()
    X = Y = 51
    u = np.zeros((X, Y, 2)); rho = np.ones((X, Y, 1))
    f = orc.equilibrium(u, rho)
    d.set_f(f)
    for _ in range(STEPS):
        orc.specular_step(f, u, rho, omega, rho_in, rho_out)
    d.step(STEPS)
    rho_o = orc.calc_rho(f); u_o = orc.calc_u(f, rho_o)
    assert fields_close(d, rho_o, u_o) < TOL


def test_gravity_10k(orc):
    Fg = (-0.0003, 0.0)
    d, omega = cases.gravity(Fg=Fg)
    X = Y = 21
    u = np.zeros((X, Y, 2)); rho = np.ones((X, Y, 1))
    f = orc.incomp_equilibrium(u, rho)
    d.set_f(f)
    for _ in range(STEPS):
        orc.gravity_step(f, u, rho, omega, 1.0, 1.0, np.array(Fg))
    d.step(STEPS)
    rho_o = orc.calc_rho(f); u_o = orc.calc_incomp_u(f) + np.array(Fg)  # the driver's u carries += Fg (gravity_test.cpp:143)
    assert fields_close(d, rho_o, u_o) < TOL


def test_free_stream_10k(orc):
    X, Y, omega, uwx = 33, 22, 1.0 / 0.8, 0.05
    d = cases.free_stream(X, Y, omega, uwx)
    u = np.zeros((X, Y, 2)); u[..., 0] = uwx
    rho = np.ones((X, Y, 1))
    f = orc.incomp_equilibrium(u, rho)
    d.set_f(f)
    for _ in range(STEPS):
        orc.free_stream_step(f, u, rho, omega, uwx)
    d.step(STEPS)
    rho_o = orc.calc_rho(f); u_o = orc.calc_incomp_u(f)
    assert fields_close(d, rho_o, u_o) < TOL


def test_cylinder_ibm_10k(orc):
    """driver 11 at a laminar Reynolds number (steady wake): the 4-iteration IBM forcing included"""
    X, Y, omega, u_lb = 96, 80, 1.0 / 0.8, 0.04   # nu = 0.1, D = 20 -> Re = 8
    th = 2 * np.pi * np.arange(64) / 64
    xs, ys = 30.3 + 10.0 * np.cos(th), 40.2 + 10.0 * np.sin(th)
    d = cases.cylinder(X, Y, omega, u_lb, xs, ys)
    ib = orc.ibm_create(xs, ys)
    u = np.zeros((X, Y, 2)); u[..., 0] = u_lb
    rho = np.ones((X, Y, 1))
    f = orc.incomp_equilibrium(u, rho)
    d.set_f(f)
    for _ in range(STEPS):
        orc.cylinder_step(f, u, rho, omega, u_lb, ib)
    d.step(STEPS)
    orc.ibm_destroy(ib)
    rho_o = orc.calc_rho(f); u_o = orc.calc_u(f, rho_o)
    assert fields_close(d, rho_o, u_o) < TOL
    assert np.isfinite(u_o).all() and float(np.abs(u_o).max()) < 0.2


def test_sedimentation_10k(orc):
    g = cases.golden("sedimentation_176x264")
    X, Y = int(g["X"]), int(g["Y"])
    omega, u_lb, w_s = float(g["omega"]), float(g["u_lb"]), float(g["w_s"])
    walls = [int(v) for v in g["walls"]]
    C_w = g["C_w"]
    d = cases.sedimentation(X, Y, omega, u_lb, w_s, C_w, walls)
    f, gg, u, rho, Cc = orc.sedimentation_init(X, Y, u_lb, C_w)
    d.set_f(f, 0)
    d.set_f(gg, 1)
    steps = STEPS
    for _ in range(steps):
        orc.sedimentation_step(f, gg, u, rho, Cc, omega, u_lb, w_s, C_w, *walls)
    d.step(steps)
    rho_g, u_g = d.get_moments(0)
    C_g, _ = d.get_moments(1)
    assert np.abs(rho_g - rho).max() < TOL and np.abs(u_g - u).max() < TOL and np.abs(C_g - Cc).max() < TOL


def mrtcg_params(R, C, Fg, add_force):
    p = MrtcgParams()
    p.R, p.C = R, C
    p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = 3.0, 0.7, 0.04, 0.7
    p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = 1.0, 0.1, 0.04, -0.7
    p.sigma, p.delta = 0.1, 0.1
    p.Fg[0], p.Fg[1] = Fg
    p.add_force = add_force
    return p


def test_mrtcg_rayleigh_taylor_long(orc):
    """driver 16 at 64 x 48.  The reference's recolouring term divides by (1e-20 + |grad phase|)
    (mrtcg_rayleigh_taylor.cpp:302-318): in the bulk, where the gradient is rounding noise, its DIRECTION is
    O(1) noise multiplied by the minority density (~1e-5 next to the top wall after ~70 steps).  The
    reference algorithm is therefore ill-conditioned: the oracle started from a state perturbed by 1e-15
    departs from itself by ~5e-7 at step 80 and ~1.6e-6 from step 1000 on (measured; it saturates, it is
    not a growing instability).  So: the 1e-9 bar is checked while the problem is still well conditioned
    (50 steps), and at 10^4 steps the CUDA path must be no further from the oracle than a few times the
    oracle's own sensitivity — which is asserted to exceed 1e-9, i.e. no implementation could do better."""
    R, C, Fg = 64, 48, (6.25e-6, 0.0)
    p = mrtcg_params(R, C, Fg, 1)
    st = orc.mrtcg_init(p, "rt")
    twin = orc.mrtcg_init(p, "rt")
    twin["r_adv"] *= 1.0 + 1e-15 * np.random.default_rng(1).standard_normal(twin["r_adv"].shape)
    d = cases.mrtcg(R, C, Fg, 1)
    d.init_two_phase(st["r_rho"], st["b_rho"], st["u"])

    def gap():
        rho, u = d.get_moments()
        ph, rr, rb = d.get_phase()
        a, b = st["r_rho"][..., 0] / 3.0, st["b_rho"][..., 0] / 1.0
        return max(np.abs(rho - st["rho"]).max(), np.abs(u - st["u"]).max(), np.abs(ph - (a - b) / (a + b)).max(),
                   np.abs(rr - st["r_rho"][..., 0]).max(), np.abs(rb - st["b_rho"][..., 0]).max())

    def advance(n):
        for _ in range(n):
            orc.mrtcg_step(p, st)
            orc.mrtcg_step(p, twin)
        d.step(n)

    advance(50)
    assert gap() < TOL
    advance(STEPS - 50)
    sensitivity = max(np.abs(twin["rho"] - st["rho"]).max(), np.abs(twin["u"] - st["u"]).max())
    assert sensitivity > 1e-7
    assert gap() < 10.0 * sensitivity
    # away from the top wall the run is still on the oracle to ~1e-8
    rho, _ = d.get_moments()
    assert np.abs(rho - st["rho"])[40:].max() < 1e-7


def test_mrtcg_static_droplet_10k(orc):
    """driver 18 at 72 x 72.  The minority density at the droplet centre follows the DIRECTION of a
    noise-level gradient (tests/golden/make_golden.py), hence 1e-8 on the colour densities there."""
    R = C = 72
    Fg = (0.0, -6.25e-6)
    p = mrtcg_params(R, C, Fg, 0)
    st = orc.mrtcg_init(p, "droplet")
    d = cases.mrtcg(R, C, Fg, 0)
    d.init_two_phase(st["r_rho"], st["b_rho"], st["u"])
    for _ in range(STEPS):
        orc.mrtcg_step(p, st)
    d.step(STEPS)
    rho, u = d.get_moments()
    _, rr, rb = d.get_phase()
    assert np.abs(rho - st["rho"]).max() < TOL and np.abs(u - st["u"]).max() < TOL
    assert np.abs(rr - st["r_rho"][..., 0]).max() < 1e-8 and np.abs(rb - st["b_rho"][..., 0]).max() < 1e-8


def test_rk_static_droplet_10k(orc):
    """driver 17 (L = 101, radius 25) and its Laplace-law reading: the pressure jump is stationary"""
    Ln = 101
    p = RkParams()
    p.L, p.radius = Ln, 25.0
    p.r_rho0, p.r_alpha, p.r_A, p.r_nu = 1.2, 1.0 / 3.0, 1e-4, 0.16
    p.b_rho0, p.b_alpha, p.b_A, p.b_nu = 1.0, 0.2, 1e-4, 0.14
    p.delta = 0.98
    st = orc.rk_init(p)
    d = cases.rk(Ln)
    d.set_f(st["r_adv"], 0)
    d.set_f(st["b_adv"], 1)
    for _ in range(STEPS):
        orc.rk_step(p, st)
    d.step(STEPS)
    rho, u = d.get_moments()
    _, rr, rb = d.get_phase()
    assert np.abs(rho[..., 0] - st["rho"]).max() < TOL and np.abs(u - st["u"]).max() < TOL
    assert np.abs(rr - st["r_rho"]).max() < TOL and np.abs(rb - st["b_rho"]).max() < TOL
    # Laplace law (SURVEY §8 note on config 4): p = sum_k rho_k (3/5)(1 - alpha_k); jump > 0 and equal on both sides
    pr = lambda a, b: a * 0.6 * (1 - 1.0 / 3.0) + b * 0.6 * (1 - 0.2)
    jump_gpu = pr(rr[50, 50], rb[50, 50]) - pr(rr[2, 2], rb[2, 2])
    jump_orc = pr(st["r_rho"][50, 50], st["b_rho"][50, 50]) - pr(st["r_rho"][2, 2], st["b_rho"][2, 2])
    assert abs(jump_gpu - jump_orc) < TOL and jump_gpu > 0.0
