#!/usr/bin/env python
"""Where do the single-pass CSF step (LBM_CSF_FUSED=1, k_csf_fused) and the three-pass step part on the device?

    python tools/csf_diag.py [--lib nofma] [R C [rpb]]

Prints, per step and per field (red / blue populations, interfacial tension, phase): max |delta| between the two steps,
how many entries differ, the first differing index; then the same kernel against itself (run twice: determinism; two band
heights: the ring and its barriers), and every variant's distance to the CPU oracle next to the oracle's distance to a twin
perturbed by 1e-15.  --lib nofma loads liblbm_b200_nofma.so (make NOFMA=1: no floating-point contraction) — if the two
steps agree bit for bit there and not in the product build, what separates them is where nvcc contracted, not a race.
TEST INFRASTRUCTURE (imports the oracle)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))

import lbm_b200 as L  # noqa: E402

args = sys.argv[1:]
if args[:1] == ["--lib"]:
    L.LIB_PATH = args[1] if "/" in args[1] else os.path.join(L.PKG_DIR, f"liblbm_b200_{args[1]}.so")
    args = args[2:]
R, C = (int(args[0]), int(args[1])) if len(args) >= 2 else (96, 64)
rpb = int(args[2]) if len(args) >= 3 else 0

import cases  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from test_gpu_csf import csf_params  # noqa: E402

NAMES = ("f_red", "f_blue", "Fs", "phase")
STEPS = 9


def run(fused, pipe="0", band=0):
    os.environ["LBM_CSF_FUSED"] = fused
    os.environ["LBM_CSF_PIPE"] = pipe
    if band:
        os.environ["LBM_TP_RPB"] = str(band)
    else:
        os.environ.pop("LBM_TP_RPB", None)
    p = csf_params(R, C)
    st = Oracle().csf_init(p)
    d = cases.csf(R, C)
    d.init_two_phase(st["r_rho"], st["b_rho"], st["u"])
    out = []
    for _ in range(STEPS):
        d.step(1)
        out.append((d.get_f(0), d.get_f(1), d.get_interfacial_tension(), d.get_phase()[0]))
    d.close()
    return out


def compare(tag, A, B):
    same = True
    for n, (a, b) in enumerate(zip(A, B), 1):
        row = []
        for name, x, y in zip(NAMES, a, b):
            diff = np.abs(x - y)
            nd = int((x != y).sum())
            if nd:
                same = False
                first = tuple(int(v) for v in np.argwhere(x != y)[0])
                row.append(f"{name}: max {diff.max():.2e} (rel {diff.max() / max(np.abs(y).max(), 1e-300):.1e}) n={nd} first={first}")
        if row:
            print(f"  [{tag}] step {n}: " + "; ".join(row))
    print(f"[{tag}] {'BIT-IDENTICAL over %d steps' % STEPS if same else 'differs'}")
    return same


print(f"library: {os.path.basename(L.LIB_PATH)}  grid {R}x{C} rpb {rpb}  {L.version()}")
three = run("0", band=rpb)
fused = run("1", band=rpb)
compare("fused vs three-pass", fused, three)
compare("fused vs fused again", run("1", band=rpb), fused)
compare("fused, 16-row bands vs default", run("1", band=16), fused)
compare("fused+pipe vs fused", run("1", pipe="1", band=rpb), fused)
compare("three-pass again", run("0", band=rpb), three)

orc = Oracle()
p = csf_params(R, C)
st, twin = orc.csf_init(p), orc.csf_init(p)
twin["r_adv"] *= 1.0 + 1e-15 * np.random.default_rng(7).standard_normal(twin["r_adv"].shape)
print("step: oracle-twin | three-pass-oracle | fused-oracle | fused-three   (max rel err over both colours)")
for n in range(STEPS):
    orc.csf_step(p, st)
    orc.csf_step(p, twin)
    own = max(cases.relerr(twin["r_adv"], st["r_adv"]), cases.relerr(twin["b_adv"], st["b_adv"]))
    e3 = max(cases.relerr(three[n][0], st["r_adv"]), cases.relerr(three[n][1], st["b_adv"]))
    e1 = max(cases.relerr(fused[n][0], st["r_adv"]), cases.relerr(fused[n][1], st["b_adv"]))
    e13 = max(cases.relerr(fused[n][0], three[n][0]), cases.relerr(fused[n][1], three[n][1]))
    print(f"  {n + 1}: {own:.2e} | {e3:.2e} | {e1:.2e} | {e13:.2e}")
