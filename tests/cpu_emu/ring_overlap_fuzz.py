"""TEST INFRASTRUCTURE (see cuda_emu.hpp): the two-phase ring step with its halo exchanges behind the interior bands
(LBM_TP_OVERLAP=1, tp_steps_ring) against the monolithic run, over random slab heights, band heights, ring sizes and
lbm_step call patterns — one THREAD per rank on the emulated device, in-process NCCL stand-in.  What it is after: the
partition of the bands into edge / interior launches (one or two edge bands at the far end, exactly one interior band, slabs
of 16 rows), the cut / rest split of the region pass, and the in-line prologue + join of every lbm_step call.
usage: ring_overlap_fuzz.py [SEED [CASES]]"""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"), os.path.join(ROOT, "tests")]
BUILD = os.path.join(HERE, "_build")
os.environ["LBM_NCCL_LIB"] = os.path.join(BUILD, "libnccl_emu.so")

import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402

L.LIB_PATH = os.path.join(BUILD, "liblbm_b200_emu.so")
import cases  # noqa: E402
from oracle_lib import MrtcgParams, Oracle, RkParams  # noqa: E402


def split_calls(rng, total):
    """total steps as a list of lbm_step arguments (1 = the in-line step, >= 2 = tp_steps_ring after the first)"""
    out = []
    while total > 0:
        n = int(min(total, rng.integers(1, 6)))
        out.append(n)
        total -= n
    return out


def one_case(rng, k):
    world = int(rng.integers(2, 5))
    model = "mrtcg" if k % 2 == 0 else "rk"
    X = int(rng.integers(16 * world, 34 * world))  # 16 .. 33 rows per rank (lbm_comm_init wants lbm_decompose_rows' split)
    rpb = int(rng.integers(4, 9))
    calls = split_calls(rng, int(rng.integers(6, 12)))
    return dict(world=world, model=model, X=X, rpb=rpb, calls=calls)


def run_case(c):
    world, X = c["world"], c["X"]
    bounds = [L.decompose_rows(X, world, r)[0] for r in range(world)] + [X]
    orc = Oracle()
    if c["model"] == "mrtcg":
        C = 36
        p = MrtcgParams()
        p.R, p.C = X, C
        p.r_rho0, p.r_alpha, p.r_nu, p.r_beta = 3.0, 0.7, 0.04, 0.7
        p.b_rho0, p.b_alpha, p.b_nu, p.b_beta = 1.0, 0.1, 0.04, -0.7
        p.sigma, p.delta = 0.1, 0.1
        p.Fg[0], p.Fg[1] = 6.25e-6, 0.0
        p.add_force = 1
        st = orc.mrtcg_init(p, "rt")
        st = {k: np.array(st[k], copy=True) for k in ("r_rho", "b_rho", "u")}
        make = lambda **slab: cases.mrtcg(X, C, (6.25e-6, 0.0), 1, **slab)  # noqa: E731
        init = lambda d, a, b: d.init_two_phase(st["r_rho"][a:b], st["b_rho"][a:b], st["u"][a:b])  # noqa: E731
    else:
        rp = RkParams()
        rp.L, rp.radius = X, X / 4.0
        rp.r_rho0, rp.r_alpha, rp.r_A, rp.r_nu = 1.2, 1.0 / 3.0, 1e-4, 0.16
        rp.b_rho0, rp.b_alpha, rp.b_A, rp.b_nu = 1.0, 0.2, 1e-4, 0.14
        rp.delta = 0.98
        rst = orc.rk_init(rp)
        rst = {k: np.array(rst[k], copy=True) for k in ("r_adv", "b_adv")}
        make = lambda **slab: cases.rk(X, **slab)  # noqa: E731

        def init(d, a, b):
            d.set_f(rst["r_adv"][a:b], 0)
            d.set_f(rst["b_adv"][a:b], 1)

    os.environ["LBM_TP_RPB"], os.environ["LBM_TP_OVERLAP"] = str(c["rpb"]), "1"
    ident = L.comm_unique_id()
    parts, errors = [None] * world, []

    def worker(rank):
        try:
            a, b = int(bounds[rank]), int(bounds[rank + 1])
            d = make(x0=a, x1=b, device=0)
            d.comm_init(ident, world, rank)
            init(d, a, b)
            for n in c["calls"]:
                d.step(n)
            parts[rank] = (d.get_f(0), d.get_f(1))
            d.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(f"rank {rank}: {type(e).__name__}: {e}")

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    os.environ.pop("LBM_TP_OVERLAP", None)
    if errors:
        return errors
    os.environ.pop("LBM_TP_RPB", None)
    mono = make(device=0)
    init(mono, 0, X)
    mono.step(sum(c["calls"]))
    want = (mono.get_f(0), mono.get_f(1))
    mono.close()
    got = tuple(np.concatenate([parts[r][l] for r in range(world)], axis=0) for l in (0, 1))
    return [] if all(np.array_equal(g, w) for g, w in zip(got, want)) else ["differs from the monolithic run"]


def main(seed, ncases):
    L.load()
    rng = np.random.default_rng(seed)
    bad = 0
    for k in range(ncases):
        c = one_case(rng, k)
        err = run_case(c)
        print(f"case {k}: {c} -> {'bit-exact' if not err else err}", flush=True)
        bad += bool(err)
    print(f"ring overlap fuzz: {ncases - bad} of {ncases} cases bit-exact vs monolithic")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 10))
