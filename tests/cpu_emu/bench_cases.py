"""TEST INFRASTRUCTURE (see cuda_emu.hpp): bench.py's per-workload plumbing — domain construction, presets, bodies, the
pinned-host initial state, import, a few steps, the moment export — run for EVERY workload at a small size against the
emulated library, so a typo in a workload the GPU box has not run yet shows up here.  Timing code is not exercised."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import lbm_b200 as L  # noqa: E402

L.LIB_PATH = os.path.join(HERE, "_build", "liblbm_b200_emu.so")
import bench  # noqa: E402


def pageable(self, shape):
    t = torch.empty(shape, dtype=torch.float64)
    return t, t.numpy()


bench.Case.pinned = pageable  # no driver here to pin host memory with

for w in sorted(bench.WORKLOADS):
    args = types.SimpleNamespace(workload=w, X=160, Y=200 if w != "rk_droplet" else 160, gpus=1)
    c = bench.Case(L, torch, args, 0, 1, 0)
    c.setup()
    assert c.import_state() > 0
    c.d.step(4)
    assert c.export_moments() > 0
    assert np.isfinite(c.rho).all() and np.isfinite(c.uo).all(), w
    assert 0.5 < float(c.rho.mean()) < 4.0, (w, float(c.rho.mean()))
    assert c.d.kernel_launches() > 0 and c.dominant_nodes() > 0
    c.d.close()
    print(w, "ok")
