"""Multi-process check of column faces bound across ranks (run under torchrun on 4 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 tests/mp_blocks_check.py

test/decompose_domain_loop.cpp's four blocks, one per rank (lbm_comm_init_blocks, lbm_link_face_rank,
lbm_comm_faces_commit, plain lbm_step), against the same blocks linked inside one process on rank 0's GPU
(lbm_link_face + lbm_step_group, which tests/test_gpu_blocks.py ties to the oracle and the reference binary): bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import lbm_b200 as L  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 4, "four blocks, four ranks"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Ln, steps, omega = 128, 65, 1.0 / 0.8
    rng = np.random.default_rng(3)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    shapes = {"A": (Ln, Ln // 4), "B": (Ln // 4, Ln // 2), "C": (Ln, Ln // 4), "D": (Ln // 4, Ln // 2)}
    f0 = {k: w * (1.0 + 0.05 * rng.standard_normal(shapes[k] + (9,))) for k in cases.LOOP_BLOCKS}  # same seed on every rank
    ident = [L.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    d = cases.loop_block_on_rank(Ln, omega, rank, ident[0], device=local)
    d.comm_faces_commit()
    d.bc_commit()
    d.set_f(f0[cases.LOOP_BLOCKS[rank]])
    d.step(steps // 2)
    d.get_f()  # an export in the middle, on every rank
    d.step(steps - steps // 2)
    got = [None] * world
    dist.all_gather_object(got, d.get_f())
    d.close()
    ok = True
    if rank == 0:
        dom = cases.loop_blocks(Ln, omega, device=local)
        for k in cases.LOOP_BLOCKS:
            dom[k].set_f(f0[k])
        L.step_group([dom[k] for k in cases.LOOP_BLOCKS], steps)
        ok = all(np.array_equal(got[i], dom[k].get_f()) for i, k in enumerate(cases.LOOP_BLOCKS))
        print(f"blocks across 4 ranks: bit-exact vs linked blocks after {steps} steps = {ok}", flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
