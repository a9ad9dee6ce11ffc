"""GPU parity of LBM_MODEL_KBC — the entropic central-moment collision ulbm::d2q9::kbc (src/ulbm.cpp) and its two
drivers — against the CPU oracle, which tests/test_oracle_vs_reference.py pins on the
compiled reference class, and against snapshots of the reference's own double-shear-flow driver."""
import os

import numpy as np
import pytest

import cases
import lbm_b200 as L
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
NU = 1.70766666e-4
S2 = 1.0 / (0.5 + 3.0 * NU)   # test/ulbm_double_shear_flow.cpp:78-79


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("kind,fresh", [(L.EQ_KBC_FRESH, True), (L.EQ_KBC, False)])
def test_initial_equilibrium(orc, kind, fresh):
    m0, u = cases.double_shear_fields(40, 56)
    d = cases.kbc(40, 56, S2)
    d.init_equilibrium(m0[..., None], u, kind)
    assert cases.relerr(d.get_f(), orc.kbc_equilibrium(m0, u, fresh_object=fresh)) < 1e-15


@pytest.mark.parametrize("R,C", [(32, 32), (37, 70), (128, 128)])
def test_double_shear_flow_vs_oracle(orc, R, C):
    m0, u = cases.double_shear_fields(R, C)
    f = orc.kbc_equilibrium(m0, u)
    d = cases.kbc(R, C, S2)
    d.set_f(f)
    d.set_moments(m0[..., None], u)   # the driver's first collide() sees the initial m0, m1 members
    n = 0
    for upto in (1, 2, 10, 100, 400):
        d.step(upto - n)
        while n < upto:
            orc.kbc_step(f, m0, u, S2)
            n += 1
        assert cases.relerr(d.get_f(), f) < 1e-12, upto
        rho, uu = d.get_moments()
        assert np.abs(rho[..., 0] - m0).max() < 1e-12 and np.abs(uu - u).max() < 1e-12, upto


def test_poiseuille_cold_start_vs_oracle(orc):
    """test/ulbm_poiseuille.cpp: adve_f = 0, m0 = 1, m1 = 0 at t = 0; pressure rows on coll_f, bounce-back columns.
    1e-12 on the first three steps.  The cold start is a violent transient that amplifies rounding (the oracle itself
    departs from a twin perturbed by 1e-15 after step 1 by 2e-11 at step 20 and 7e-13 at step 300), so from then on the
    bound is the algorithm's own conditioning: no further from the oracle than 5 x the oracle is from that twin."""
    R, C = 48, 40
    nu = 1e-4
    s2 = 1.0 / (0.5 + 3.0 * nu)
    rho_out = 1.0
    rho_in = 3.0 * (R - 1) * (8.0 * nu * 0.05 / (C * C)) + rho_out
    f = np.zeros((R, C, 9)); m0 = np.ones((R, C)); u = np.zeros((R, C, 2))
    tf, tm0, tu = f.copy(), m0.copy(), u.copy()
    d = cases.kbc(R, C, s2, poiseuille=(rho_in, rho_out))
    d.set_f(f)
    d.set_moments(m0[..., None], u)
    done = 0
    for upto in (1, 2, 3, 20, 300):
        for n in range(done, upto):
            orc.kbc_step(f, m0, u, s2, 1, rho_in, rho_out)
            orc.kbc_step(tf, tm0, tu, s2, 1, rho_in, rho_out)
            if n == 0:
                tf *= 1.0 + 1e-15 * np.random.default_rng(7).standard_normal(tf.shape)
        d.step(upto - done)
        done = upto
        own = cases.relerr(tf, f)
        tol = 1e-12 if upto <= 3 else max(1e-12, 5.0 * own)
        assert tol < 1e-9 and cases.relerr(d.get_f(), f) < tol, (upto, own)
    rho, uu = d.get_moments()
    assert np.abs(rho[..., 0] - m0).max() < tol and np.abs(uu - u).max() < tol


def test_without_set_moments_the_populations_decide(orc):
    R, C = 24, 30
    m0, u = cases.double_shear_fields(R, C)
    f = orc.kbc_equilibrium(m0, u, fresh_object=False)
    d = cases.kbc(R, C, S2)
    d.set_f(f)
    rho_o = orc.calc_rho(f); u_o = orc.calc_u(f, rho_o)
    m0b, ub = rho_o[..., 0].copy(), u_o.copy()
    orc.kbc_step(f, m0b, ub, S2)
    d.step(1)
    assert cases.relerr(d.get_f(), f) < 1e-12


def test_slabs_and_graph_equal_monolithic(orc):
    R, C = 45, 64
    m0, u = cases.double_shear_fields(R, C)
    f = orc.kbc_equilibrium(m0, u)
    mono = cases.kbc(R, C, S2)
    mono.set_f(f); mono.set_moments(m0[..., None], u)
    g = cases.kbc(R, C, S2)
    g.use_graph(True)
    g.set_f(f); g.set_moments(m0[..., None], u)
    slabs = []
    for r in range(3):
        x0, x1 = L.decompose_rows(R, 3, r)
        s = cases.kbc(R, C, S2, x0=x0, x1=x1)
        slabs.append(s)
    for r, s in enumerate(slabs):
        s.link(slabs[(r - 1) % 3], slabs[(r + 1) % 3])
        s.set_f(f[s.cfg.x0:s.cfg.x1]); s.set_moments(m0[s.cfg.x0:s.cfg.x1, :, None], u[s.cfg.x0:s.cfg.x1])
    for n in (1, 6, 31):
        mono.step(n); g.step(n); L.step_group(slabs, n)
        want = mono.get_f()
        assert np.array_equal(g.get_f(), want)
        assert np.array_equal(np.concatenate([s.get_f() for s in slabs], axis=0), want)


@pytest.mark.skipif(not os.path.exists(os.path.join(cases.GOLDEN, "kbc_double_shear_128.npz")), reason="golden not generated")
def test_double_shear_flow_reference_driver_golden():
    """snapshots written by the reference's own test/ulbm_double_shear_flow.cpp (128 x 128, its hard-wired size)"""
    g = cases.golden("kbc_double_shear_128")
    R = C = 128
    m0, u = cases.double_shear_fields(R, C)
    d = cases.kbc(R, C, S2)
    d.init_equilibrium(m0[..., None], u, L.EQ_KBC_FRESH)   # kbc.eval_equilibrium(kbc.adve_f) on the fresh object (:97)
    d.set_moments(m0[..., None], u)
    t = 0
    for k, s in enumerate(int(s) for s in g["steps"]):
        d.step(s - t)
        t = s
        rho, uu = d.get_moments()
        tol = float(g["tol"][k])
        st = int(g["stride"])
        assert np.abs(uu[::st, ::st, 0] - g["ux"][k]).max() < tol, s
        assert np.abs(uu[::st, ::st, 1] - g["uy"][k]).max() < tol, s
        assert np.abs(rho[::st, ::st, 0] - g["rho"][k]).max() < tol, s


@pytest.mark.parametrize("R,C", [(3, 3), (4, 5), (6, 131), (130, 7)])
def test_tiny_and_ragged_grids(orc, R, C):
    rng = np.random.default_rng(R * 100 + C)
    m0 = 1.0 + 0.01 * rng.standard_normal((R, C))
    u = 0.02 * rng.standard_normal((R, C, 2))
    f = orc.kbc_equilibrium(m0, u, fresh_object=False)
    for poiseuille in (None, (1.0003, 1.0)) if R >= 4 else (None,):
        fo, a0, a1 = f.copy(), m0.copy(), u.copy()
        d = cases.kbc(R, C, S2, poiseuille=poiseuille)
        d.set_f(fo)
        d.set_moments(a0[..., None], a1)
        for _ in range(5):
            orc.kbc_step(fo, a0, a1, S2, 0 if poiseuille is None else 1, *(poiseuille or (1.0, 1.0)))
        d.step(5)
        assert cases.relerr(d.get_f(), fo) < 1e-12, poiseuille
