"""GPU parity tests of what was written AFTER this round's GPU minutes were spent: they have passed on the emulated device
(tests/cpu_emu) but not yet on hardware.  The file sorts last on purpose: under `pytest -x` a first-run
surprise here cannot hide the result of the suite that has run on the B200 before.  Once they have passed on the device
they move next to their families (test_gpu_bgk.py, test_gpu_two_phase.py, test_gpu_snapshot.py).
  - BGK + ADE with an immersed body (force = LBM_FORCE_IBM on LBM_MODEL_BGK_ADE)
  - lbm_rk_diagnostics and the RK driver's nineteen output files"""
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
import lbm_b200 as L
from oracle_lib import Oracle
from test_gpu_two_phase import rk_params

# first runs on hardware: a kernel that hangs there must end this process, not the box's whole GPU session
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_STEP = 1e-12


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_sedimentation_with_immersed_body_vs_oracle(orc):
    """BASELINE configs[4] as worded ("with immersed-boundary coupling"): driver 15's two lattices plus a body coupled the
    way test/cylinder_test.cpp couples one.  Not a reference driver; the oracle composes its two pinned steps."""
    X, Y = 96, 120
    omega, u_lb, w_s = 1.0 / 0.8, 0.02, 3e-3
    C_w = np.zeros(X); C_w[-20:] = 1e-3
    walls = (-30, 70, 90)
    th = 2 * np.pi * np.arange(40) / 40
    xs, ys = 30.3 + 7.2 * np.cos(th), 35.6 + 7.2 * np.sin(th)
    d = cases.sedimentation(X, Y, omega, u_lb, w_s, C_w, walls, ibm=True)
    d.ibm_set_markers(xs, ys)
    f, gg, u, rho, Cc = orc.sedimentation_init(X, Y, u_lb, C_w)
    d.set_f(f, 0)
    d.set_f(gg, 1)
    ib = orc.ibm_create(xs, ys)
    plain_f = f.copy()
    for n in range(1, 41):
        orc.sedimentation_step(f, gg, u, rho, Cc, omega, u_lb, w_s, C_w, *walls, ib=ib)
        d.step(1)
        if n in (1, 2, 10, 40):
            assert cases.relerr(d.get_f(0), f) < TOL_STEP and cases.relerr(d.get_f(1), gg) < TOL_STEP, n
    orc.ibm_destroy(ib)
    # the body is felt: the state differs from the run without it
    pf, pg, pu, prho, pC = orc.sedimentation_init(X, Y, u_lb, C_w)
    for _ in range(40):
        orc.sedimentation_step(pf, pg, pu, prho, pC, omega, u_lb, w_s, C_w, *walls)
    assert np.abs(pf - f).max() > 1e-6


def test_rk_diagnostic_fields_vs_oracle_and_golden(orc):
    """lbm_rk_diagnostics: the fields driver 17 snapshots at the top of every iteration (normal with the 0.1 max|grad|
    cut, curvature, interfacial tension, eta, kappa, 1/tau, the red colour's omega1/2/3), against the oracle at every
    step and against the reference driver's own nx / ny / ks / norms / fx / fy / kappas / omegas1 / omegas2 files"""
    from test_oracle_golden import rk_diag_fields

    g = cases.golden("rk_droplet_101")
    p = rk_params(101)
    st = orc.rk_init(p)
    d = cases.rk(101)
    d.set_f(st["r_adv"], 0)
    d.set_f(st["b_adv"], 1)
    steps = [int(v) for v in g["steps"]]
    for n in range(steps[-2] + 1):
        want = orc.rk_diagnostics(p, st)
        got = d.rk_diagnostics(5e-3)
        scale = {k: max(float(np.abs(v).max()), 1e-30) for k, v in want.items()}
        for k in ("phase", "grad", "norm", "n", "K", "Fs", "eta", "kappa", "rparams", "omega1", "omega2", "omega3"):
            # absolute 1e-12 on O(1) fields, relative on the small ones (Fs ~ 1e-5, omega2 ~ 1e-6)
            assert np.abs(got[k] - want[k]).max() < 1e-12 * max(scale[k], 1e-3), (n, k)
        if n in steps:
            for name, a in rk_diag_fields(got).items():
                assert np.abs(a - g[name][steps.index(n)]).max() < 1e-12, (n, name)
        orc.rk_step(p, st)
        d.step(1)
    assert cases.relerr(d.get_f(0), st["r_adv"]) < 1e-12  # asking for diagnostics between steps does not disturb the state


def test_rk_diagnostics_error_behaviour():
    d = cases.mrtcg(16, 16, (0.0, 0.0), 0)
    with pytest.raises(L.LbmError, match="LBM_MODEL_RK"):
        d.rk_diagnostics()
    slab = cases.rk(32, x0=0, x1=16)
    slab.init_two_phase(np.ones((16, 32)), np.ones((16, 32)), np.zeros((16, 32, 2)))
    with pytest.raises(L.LbmError, match="monolithic domains or the ranks"):
        slab.rk_diagnostics()  # a slab outside a ring: no way to reduce max|grad| or to swap the normal's halo


def test_rk_driver_writes_all_nineteen_reference_files(tmp_path):
    """drivers/rk_static_droplet mirrors test/rk_static_droplet_test.cpp:617-635 file for file; the diagnostic stacks
    (normal, curvature, interfacial tension, kappa, omega1/2/3) equal the reference driver's own output"""
    exe = os.path.join(ROOT, "drivers", "bin", "rk_static_droplet")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "drivers")], stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, "101", "12"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    g = cases.golden("rk_droplet_101")
    names = {"r-fs": "r_fs", "b-fs": "b_fs", "ux": "ux", "uy": "uy", "nx": "nx", "ny": "ny", "rho": "rho", "rhon": "rhon", "ks": "ks",
             "norms": "norms", "fx": "fx", "fy": "fy", "gradx": "gradx", "grady": "grady", "rparams": "rparams", "kappas": "kappas",
             "omegas1": "omegas1", "omegas2": "omegas2", "omegas3": None}
    got = {}
    for fname, key in names.items():
        path = tmp_path / f"rk-static-droplet-{fname}.pt"
        assert path.exists(), fname
        a = list(torch.jit.load(str(path)).parameters())[0].numpy()
        assert a.shape[:2] == (101, 101) and a.shape[-1] == 12, (fname, a.shape)
        got[fname] = a
        if key is not None:
            for k, s in enumerate(int(s) for s in g["steps"]):
                if s < 12:
                    assert np.abs(a[..., s] - g[key][k]).max() < 1e-12, (fname, s)
    assert np.abs(got["omegas3"] - (got["omegas1"] + got["omegas2"])).max() < 1e-15


@pytest.mark.parametrize("pipe", ["0", "1"])
@pytest.mark.parametrize("R,C,rpb", [(96, 64, 0), (41, 33, 0), (70, 300, 0), (200, 131, 16), (130, 125, 128), (16, 12, 0)])
def test_csf_single_pass_equals_three_pass(monkeypatch, R, C, rpb, pipe):
    """LBM_CSF_FUSED=1 (k_csf_fused: moments -> ring, normals at lag 2, collision at lag 5, Fs double-buffered) against the
    three-pass step: populations of both colours and the interfacial tension bit for bit, several strips and bands, a first
    step out of an import (three-pass) followed by fused steps, getters in between."""
    from test_gpu_csf import csf_params

    p = csf_params(R, C)
    st = Oracle().csf_init(p)
    runs = []
    for fused in ("0", "1"):
        monkeypatch.setenv("LBM_CSF_FUSED", fused)
        monkeypatch.setenv("LBM_CSF_PIPE", pipe)  # the software-pipelined variant of the kernel (two resident blocks)
        if rpb:
            monkeypatch.setenv("LBM_TP_RPB", str(rpb))
        d = cases.csf(R, C)
        d.init_two_phase(st["r_rho"], st["b_rho"], st["u"])
        out = []
        for n in (1, 1, 3, 4):
            d.step(n)
            out.append((d.get_f(0), d.get_f(1), d.get_interfacial_tension(), d.get_phase()[0]))
        runs.append(out)
        d.close()
    for a, b in zip(*runs):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("P", [2, 3])
def test_csf_single_pass_on_linked_slabs(monkeypatch, orc, P):
    """the three stages of the single-pass step interleaved across linked slabs (lbm_step_group): bit-identical to the
    monolithic run, which test_csf_single_pass_equals_three_pass ties to the three-pass step"""
    import test_gpu_csf

    monkeypatch.setenv("LBM_CSF_FUSED", "1")
    test_gpu_csf.test_csf_linked_slabs_equal_monolithic(orc, P)
