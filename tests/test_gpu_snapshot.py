"""GPU: lbm_snapshot_async — fields staged on the domain's stream, copied on a copy stream under the
following steps — returns exactly what lbm_get_moments / lbm_get_phase return at the same step, and the
drivers' .pt snapshot stacks load back through torch."""
import os
import subprocess

import numpy as np
import pytest
import torch

import cases
import lbm_b200 as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W9 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def pinned(shape):
    t = torch.empty(shape, dtype=torch.float64, pin_memory=torch.cuda.is_available())  # pageable under tests/cpu_emu
    return t, t.numpy()


def test_async_snapshots_equal_synchronous_ones_bgk():
    X, Y = 200, 300
    f0 = W9 * (1.0 + 0.02 * np.random.default_rng(4).standard_normal((X, Y, 9)))
    a, b = cases.free_stream(X, Y, 1.6, 0.05), cases.free_stream(X, Y, 1.6, 0.05)
    a.set_f(f0); b.set_f(f0)
    keep = [pinned((X, Y, 1)) + pinned((X, Y, 2)) for _ in range(3)]
    want = []
    for k in range(3):
        a.step(7); b.step(7)
        want.append(b.get_moments())
        a.snapshot_async(rho=keep[k][1], u=keep[k][3])   # back-to-back snapshots: the second waits for the first copy on the device
        a.step(5); b.step(5)
    a.snapshot_wait()
    for k in range(3):
        assert np.array_equal(keep[k][1], want[k][0]) and np.array_equal(keep[k][3], want[k][1]), k
    assert np.array_equal(a.get_f(), b.get_f())


def test_async_snapshot_two_phase_with_phase():
    R, C = 90, 140
    rr = np.where(np.arange(R)[:, None] < R / 2 + 5 * np.sin(np.arange(C)[None, :] / 11.0), 3.0, 0.0)
    rb = np.where(rr > 0, 0.0, 1.0)
    u = np.zeros((R, C, 2))
    a, b = cases.mrtcg(R, C, (6.25e-6, 0.0), 1), cases.mrtcg(R, C, (6.25e-6, 0.0), 1)
    a.init_two_phase(rr, rb, u); b.init_two_phase(rr, rb, u)
    a.step(9); b.step(9)
    (_, rho), (_, uu), (_, ph) = pinned((R, C, 1)), pinned((R, C, 2)), pinned((R, C))
    a.snapshot_async(rho=rho, u=uu, phase=ph)
    a.step(4); b_rho, b_u = b.get_moments(); b_ph = b.get_phase()[0]; b.step(4)
    a.snapshot_wait()
    assert np.array_equal(rho, b_rho) and np.array_equal(uu, b_u) and np.array_equal(ph, b_ph)
    assert np.array_equal(a.get_f(0), b.get_f(0)) and np.array_equal(a.get_f(1), b.get_f(1))


def test_poiseuille_driver_writes_reference_pt_files(tmp_path):
    """drivers/horizontal_poiseuille mirrors test/horizontal_poiseuille_test.cpp down to the files it saves"""
    exe = os.path.join(ROOT, "drivers", "bin", "horizontal_poiseuille")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "drivers")], stdout=subprocess.DEVNULL)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    g = cases.golden("poiseuille_21x21")
    for name in ("hpt-ux.pt", "hpt-uy.pt", "hpt-fs.pt", "hpt-ps.pt"):
        assert (tmp_path / name).exists(), name
    ux = list(torch.jit.load(str(tmp_path / "hpt-ux.pt")).parameters())[0].numpy()
    assert ux.shape[:2] == (21, 21) and np.isfinite(ux).all()
    assert "L2" in r.stdout or "l2" in r.stdout
