"""GPU: lbm_use_graph — a steady-state step pair captured into a CUDA graph and replayed — must give
bit-identical states to plain stepping, for every model, odd and even step counts, and across imports."""
import numpy as np
import pytest

import cases
import lbm_b200 as L
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

W9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def noisy(X, Y, seed):
    return W9 * (1.0 + 0.02 * np.random.default_rng(seed).standard_normal((X, Y, 9)))


def run_both(make, load, schedule, nlat=1):
    a, b = make(), make()
    b.use_graph(True)
    load(a); load(b)
    for n in schedule:
        a.step(n); b.step(n)
        for l in range(nlat):
            assert np.array_equal(a.get_f(l), b.get_f(l)), (n, l)
    return a, b


def test_poiseuille_graph_equals_plain():
    f0 = noisy(21, 21, 1)
    run_both(lambda: cases.poiseuille()[0], lambda d: d.set_f(f0), [1, 4, 7, 100, 2, 33])


def test_cylinder_ibm_graph_equals_plain():
    g = cases.golden("cylinder_99x77")
    mk = lambda: cases.cylinder(int(g["X"]), int(g["Y"]), float(g["omega"]), float(g["u_lb"]), g["marker_x"], g["marker_y"])
    a, b = run_both(mk, lambda d: d.set_f(g["f0"]), [5, 40, 9])
    assert np.array_equal(a.ibm_get_force(), b.ibm_get_force())
    # a new import in the middle of a graph-stepped run
    f1 = a.get_f()
    a.set_f(f1); b.set_f(f1)
    a.step(11); b.step(11)
    assert np.array_equal(a.get_f(), b.get_f())


def test_sedimentation_graph_equals_plain():
    g = cases.golden("sedimentation_176x264")
    X, Y = int(g["X"]), int(g["Y"])
    f, gg, *_ = Oracle().sedimentation_init(X, Y, float(g["u_lb"]), g["C_w"])
    mk = lambda: cases.sedimentation(X, Y, float(g["omega"]), float(g["u_lb"]), float(g["w_s"]), g["C_w"], g["walls"])

    def load(d):
        d.set_f(f, 0); d.set_f(gg, 1)

    run_both(mk, load, [6, 21], nlat=2)


@pytest.mark.parametrize("model", ["mrtcg", "rk"])
def test_two_phase_graph_equals_plain(model):
    if model == "mrtcg":
        R, C = 70, 150
        mk = lambda: cases.mrtcg(R, C, (6.25e-6, 0.0), 1)
        rr = np.where(np.arange(R)[:, None] < R / 2 + 4 * np.cos(np.arange(C)[None, :] / 9.0), 3.0, 0.0)
        rb = np.where(rr > 0, 0.0, 1.0)
    else:
        R = C = 90
        mk = lambda: cases.rk(R)
        s = np.hypot(np.arange(R)[:, None] - R / 2, np.arange(C)[None, :] - C / 2)
        sg = 1.0 / (1.0 + np.exp(-2.0 * (s - 20.0)))
        rr, rb = 1.2 * (1 - sg), 1.0 * sg
    u = np.zeros((R, C, 2))
    a, b = run_both(mk, lambda d: d.init_two_phase(rr, rb, u), [1, 6, 25, 3], nlat=2)
    assert np.array_equal(a.get_phase()[0], b.get_phase()[0])
