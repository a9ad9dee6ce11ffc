"""GPU: slab decomposition along axis 0 (test/decompose_domain.cpp generalised).  Several slabs are
linked inside one process — on one GPU when only one is visible — and stepped in lock step; the
result must equal the monolithic run BIT FOR BIT (same arithmetic per node, only the indexing and
the ghost-row exchange differ), and the reference's own two-domain driver must be reproduced."""
import numpy as np
import pytest

import cases
import lbm_b200 as L
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def n_devices():
    import torch

    return torch.cuda.device_count()


def make_slabs(cfg_kw, X, P, setup, ring=True, devices=None):
    slabs = []
    for r in range(P):
        x0, x1 = L.decompose_rows(X, P, r)
        dev = devices[r] if devices else 0
        d = L.Domain(L.default_config(X=X, x0=x0, x1=x1, device=dev, **cfg_kw))
        setup(d)
        slabs.append(d)
    for r, d in enumerate(slabs):
        lo = slabs[(r - 1) % P] if (ring or r > 0) else None
        hi = slabs[(r + 1) % P] if (ring or r < P - 1) else None
        d.link(lo, hi)
    return slabs


def scatter(slabs, f, lattice=0):
    for d in slabs:
        d.set_f(f[d.cfg.x0:d.cfg.x1], lattice)


def gather(slabs, lattice=0):
    return np.concatenate([d.get_f(lattice) for d in slabs], axis=0)


def test_reference_two_domain_driver_golden():
    """test/decompose_domain.cpp: A over B, each 21x21 with its own periodic wrap on the outer rows."""
    g = cases.golden("decompose_2x21x21")
    omega, rho_in, rho_out = float(g["omega"]), float(g["rho_in"]), float(g["rho_out"])
    X, Y = 42, 21
    kw = dict(model=L.MODEL_BGK, Y=Y, omega=omega, equilibrium=L.EQ_COMPRESSIBLE)
    A = L.Domain(L.default_config(X=X, x0=0, x1=21, **kw))
    B = L.Domain(L.default_config(X=X, x0=21, x1=42, **kw))
    for d in (A, B):
        d.preset_poiseuille(rho_in, rho_out)
    # the driver's advect wraps INSIDE each domain (A row 0 <- A row -1), the bind joins A[-1] and B[0]
    A.link(A, B)
    B.link(A, B)
    A.set_f(g["fA"][0]); B.set_f(g["fB"][0])
    t = 0
    for k, s in enumerate(int(s) for s in g["steps"]):
        L.step_group([A, B], s - t)
        t = s
        tol = 1e-12 if s <= 2 else 1e-9
        assert cases.relerr(A.get_f(), g["fA"][k]) < tol, s
        assert cases.relerr(B.get_f(), g["fB"][k]) < tol, s


@pytest.mark.parametrize("P", [2, 3, 5])
def test_poiseuille_slabs_equal_monolithic(P):
    X, Y = 23, 21
    omega, rho_in, rho_out = cases.channel_constants(X, Y, 0.1)
    kw = dict(model=L.MODEL_BGK, Y=Y, omega=omega, equilibrium=L.EQ_INCOMPRESSIBLE)
    mono = L.Domain(L.default_config(X=X, **kw))
    mono.preset_poiseuille(rho_in, rho_out)
    slabs = make_slabs(kw, X, P, lambda d: d.preset_poiseuille(rho_in, rho_out))
    rng = np.random.default_rng(11)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f0 = w * (1.0 + 0.02 * rng.standard_normal((X, Y, 9)))
    mono.set_f(f0); scatter(slabs, f0)
    for n in (1, 1, 7, 40):
        mono.step(n); L.step_group(slabs, n)
        assert np.array_equal(gather(slabs), mono.get_f())
    masks = np.concatenate([d.bc_mask() for d in slabs], axis=0)
    assert np.array_equal(masks, mono.bc_mask())  # decomposition indexing of the rule masks: bit-exact


@pytest.mark.parametrize("P", [2, 4])
def test_cylinder_slabs_equal_monolithic(P):
    g = cases.golden("cylinder_99x77")
    X, Y = int(g["X"]), int(g["Y"])
    omega, u_lb = float(g["omega"]), float(g["u_lb"])
    kw = dict(model=L.MODEL_BGK, Y=Y, omega=omega, equilibrium=L.EQ_COMPRESSIBLE, force=L.FORCE_IBM)
    # shift the body into the first slab
    xs, ys = g["marker_x"] - 25.0, g["marker_y"]
    if P == 4:
        xs = 12.3 + (xs - xs.mean()) * 0.2
        ys = 40.1 + (ys - ys.mean()) * 0.2
    mono = L.Domain(L.default_config(X=X, **kw))
    mono.preset_free_stream(u_lb, 0.0)
    mono.ibm_set_markers(xs, ys)
    roi = mono.ibm_roi()

    def setup(d):
        d.preset_free_stream(u_lb, 0.0)
        if d.cfg.x0 <= roi[0] and roi[1] <= d.cfg.x1:
            d.ibm_set_markers(xs, ys)

    slabs = make_slabs(kw, X, P, setup)
    assert sum(1 for d in slabs if d.cfg.x0 <= roi[0] and roi[1] <= d.cfg.x1) == 1
    mono.set_f(g["f0"]); scatter(slabs, g["f0"])
    for n in (1, 2, 30):
        mono.step(n); L.step_group(slabs, n)
        assert np.array_equal(gather(slabs), mono.get_f())


@pytest.mark.parametrize("P,shift", [(2, 0.0), (2, 3.0), (3, 0.0), (4, -11.0), (5, 0.0)])
def test_cylinder_across_slab_cuts_equal_monolithic(P, shift):
    """the ROI of the immersed body straddles one or several cuts.  Every slab that owns ROI rows keeps the
    whole solve, the slabs swap the moments of their own active nodes, and the force field — hence the state — is the
    monolithic one bit for bit.  Every slab is handed the same marker list; slabs away from the body ignore it."""
    g = cases.golden("cylinder_99x77")
    X, Y = int(g["X"]), int(g["Y"])
    omega, u_lb = float(g["omega"]), float(g["u_lb"])
    kw = dict(model=L.MODEL_BGK, Y=Y, omega=omega, equilibrium=L.EQ_COMPRESSIBLE, force=L.FORCE_IBM)
    xs, ys = g["marker_x"] + shift, g["marker_y"]
    mono = L.Domain(L.default_config(X=X, **kw))
    mono.preset_free_stream(u_lb, 0.0)
    mono.ibm_set_markers(xs, ys)
    roi = mono.ibm_roi()

    def setup(d):
        d.preset_free_stream(u_lb, 0.0)
        d.ibm_set_markers(xs, ys)

    slabs = make_slabs(kw, X, P, setup)
    owners = [d for d in slabs if d.cfg.x0 < roi[1] and roi[0] < d.cfg.x1]
    assert len(owners) >= 2, "the body must cross a cut in this test"
    mono.set_f(g["f0"]); scatter(slabs, g["f0"])
    for n in (1, 2, 30):
        mono.step(n); L.step_group(slabs, n)
        assert np.array_equal(gather(slabs), mono.get_f())
        for d in owners:
            assert np.array_equal(d.ibm_get_force(), mono.ibm_get_force())
    # an import in the middle of a run (the un-prepared-state path) and on
    f = mono.get_f()
    mono.set_f(f); scatter(slabs, f)
    mono.step(3); L.step_group(slabs, 3)
    assert np.array_equal(gather(slabs), mono.get_f())


def test_sedimentation_slabs_equal_monolithic():
    g = cases.golden("sedimentation_176x264")
    X, Y = int(g["X"]), int(g["Y"])
    omega, u_lb, w_s = float(g["omega"]), float(g["u_lb"]), float(g["w_s"])
    orc = Oracle()
    f, gg, u, rho, Cc = orc.sedimentation_init(X, Y, u_lb, g["C_w"])
    kw = dict(model=L.MODEL_BGK_ADE, Y=Y, omega=omega, omega_g=omega, equilibrium=L.EQ_COMPRESSIBLE, w_s=w_s)
    walls = [int(v) for v in g["walls"]]
    mono = L.Domain(L.default_config(X=X, **kw))
    mono.preset_sedimentation(u_lb, g["C_w"], *walls)
    slabs = make_slabs(kw, X, 3, lambda d: d.preset_sedimentation(u_lb, g["C_w"], *walls))
    for lat, a in ((0, f), (1, gg)):
        mono.set_f(a, lat); scatter(slabs, a, lat)
    for n in (1, 9):
        mono.step(n); L.step_group(slabs, n)
        assert np.array_equal(gather(slabs, 0), mono.get_f(0)) and np.array_equal(gather(slabs, 1), mono.get_f(1))


@pytest.mark.skipif("n_devices() < 2")
def test_slabs_on_two_devices_equal_monolithic():
    X, Y = 64, 96
    omega = 1.7
    kw = dict(model=L.MODEL_BGK, Y=Y, omega=omega, equilibrium=L.EQ_COMPRESSIBLE)
    mono = L.Domain(L.default_config(X=X, **kw))
    mono.preset_free_stream(0.05, 0.0)
    slabs = make_slabs(kw, X, 2, lambda d: d.preset_free_stream(0.05, 0.0), devices=[0, 1])
    rng = np.random.default_rng(3)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f0 = w * (1.0 + 0.02 * rng.standard_normal((X, Y, 9)))
    mono.set_f(f0); scatter(slabs, f0)
    mono.step(25); L.step_group(slabs, 25)
    assert np.array_equal(gather(slabs), mono.get_f())


@pytest.mark.parametrize("model,P", [("mrtcg", 2), ("mrtcg", 3), ("rk", 2), ("rk", 4)])
def test_two_phase_slabs_equal_monolithic(model, P):
    """linked two-phase slabs: population ghost rows + two-row halos of the moment planes across every cut
    (the global edge replicates); the gathered state must equal the monolithic run BIT FOR BIT"""
    if model == "mrtcg":
        R, C = 61, 140
        mk = lambda **slab: cases.mrtcg(R, C, (6.25e-6, 0.0), 1, **slab)
        rr = np.where(np.arange(R)[:, None] < R / 2 + 6 * np.cos(np.arange(C)[None, :] / 13.0), 3.0, 0.0)
        rb = np.where(rr > 0, 0.0, 1.0)
    else:
        R = C = 72
        mk = lambda **slab: cases.rk(R, **slab)
        s = np.hypot(np.arange(R)[:, None] - R / 2, np.arange(C)[None, :] - C / 2)
        sg = 1.0 / (1.0 + np.exp(-2.0 * (s - 17.0)))
        rr, rb = 1.2 * (1 - sg), 1.0 * sg
    u = 1e-4 * np.random.default_rng(2).standard_normal((R, C, 2))
    mono = mk()
    mono.init_two_phase(rr, rb, u)
    slabs = []
    for r in range(P):
        x0, x1 = L.decompose_rows(R, P, r)
        slabs.append(mk(x0=x0, x1=x1))
    for r, d in enumerate(slabs):
        d.link(slabs[(r - 1) % P], slabs[(r + 1) % P])
        d.init_two_phase(rr[d.cfg.x0:d.cfg.x1], rb[d.cfg.x0:d.cfg.x1], u[d.cfg.x0:d.cfg.x1])
    for n in (1, 1, 5, 20):
        mono.step(n); L.step_group(slabs, n)
        for lat in (0, 1):
            assert np.array_equal(gather(slabs, lat), mono.get_f(lat)), (n, lat)
    ph = np.concatenate([d.get_phase()[0] for d in slabs], axis=0)
    assert np.array_equal(ph, mono.get_phase()[0])
    with pytest.raises(L.LbmError):
        slabs[0].step(1)   # a linked slab is advanced with the group
