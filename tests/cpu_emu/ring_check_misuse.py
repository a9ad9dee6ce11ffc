"""TEST INFRASTRUCTURE (see cuda_emu.hpp): lbm_comm_check on a ring of threads over the NCCL stand-in.  An immersed body
whose ROI rows cross a slab cut, handed to rank 0 only — the misuse that hung an eight-GPU run — must come back as
LBM_ERR_COMM on EVERY rank; the two correct usages (every rank gets the list; rank 0 alone when the ROI lies inside its
slab) must pass.          usage: ring_check_misuse.py WORLD"""
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"), os.path.join(ROOT, "tests")]
BUILD = os.path.join(HERE, "_build_asan" if os.environ.get("LBM_EMU_ASAN") == "1" else "_build")  # SAN=1 build: run with libasan preloaded
os.environ["LBM_NCCL_LIB"] = os.path.join(BUILD, "libnccl_emu.so")

import numpy as np  # noqa: E402

import lbm_b200 as L  # noqa: E402

L.LIB_PATH = os.path.join(BUILD, "liblbm_b200_emu.so")


def ring(world, body_rows, who_gets_markers):
    """-> per-rank outcome of comm_check: 'ok' or the error text"""
    X, Y = 16 * world, 48
    th = 2 * np.pi * np.arange(24) / 24
    xs, ys = body_rows + 4.1 * np.cos(th), 24.3 + 4.1 * np.sin(th)
    sync = threading.Barrier(world)
    shared = {}
    out = [None] * world

    def worker(rank):
        x0, x1 = L.decompose_rows(X, world, rank)
        d = L.Domain(L.default_config(model=L.MODEL_BGK, X=X, Y=Y, omega=1.2, equilibrium=L.EQ_COMPRESSIBLE, force=L.FORCE_IBM,
                                      x0=x0, x1=x1, device=0))
        if rank == 0:
            shared["id"] = L.comm_unique_id()
        sync.wait()
        d.comm_init(shared["id"], world, rank)
        d.preset_free_stream(0.05, 0.0)
        if rank in who_gets_markers:
            d.ibm_set_markers(xs, ys)
        try:
            d.comm_check()
            out[rank] = "ok"
        except L.LbmError as e:
            out[rank] = str(e)
        sync.wait()
        d.close()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def main(world):
    L.load()
    everyone = set(range(world))
    # body across the first cut (rows 16 +- 6), every rank has the list
    assert ring(world, 16.2, everyone) == ["ok"] * world
    # body inside rank 0's slab (rows 8 +- 6 with the +-2 ROI margin: rows 1..15), rank 0 alone has the list
    assert ring(world, 8.2, {0}) == ["ok"] * world
    # body across the first cut, rank 0 alone has the list: an error on every rank, naming the slab that lacks it
    got = ring(world, 16.2, {0})
    assert all("crosses the slab of rank 1, which was given no markers" in g for g in got), got
    # ... and nobody has stepped, so nobody hangs
    print("comm_check: misuse reported on all", world, "ranks:", got[world - 1][:160])
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 2))
