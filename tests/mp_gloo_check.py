"""World-size-2 check of the host-side multi-GPU logic on CPU (gloo): launched by tests/test_mp_gloo.py as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/mp_gloo_check.py

What runs without a GPU: the rendezvous pattern bench.py uses (unique id broadcast from rank 0), the slab
decomposition every rank derives on its own, the per-slab synthetic initial states (gathered == monolithic),
and the library's refusal to compute without CUDA on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lattice-boltzmann-method_b200", "python"))
import bench  # noqa: E402
import lbm_b200 as L  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    fails = []

    # 1. the 128-byte communicator id travels from rank 0 to everyone (bench.py / tests/mp_nccl_check.py pattern)
    ident = [bytes((7 * i + 3) % 256 for i in range(L.UNIQUE_ID_BYTES)) if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    if ident[0] != bytes((7 * i + 3) % 256 for i in range(L.UNIQUE_ID_BYTES)):
        fails.append("unique id broadcast")

    # 2. every rank derives the same decomposition; the slabs tile the grid in rank order
    for X in (2 * world, 37, 8192 * world, 16384 * world + 5):
        x0, x1 = L.decompose_rows(X, world, rank)
        spans = [None] * world
        dist.all_gather_object(spans, (x0, x1))
        if spans[0][0] != 0 or spans[-1][1] != X or any(spans[i][1] != spans[i + 1][0] for i in range(world - 1)):
            fails.append(f"decompose {X}: {spans}")
        if max(b - a for a, b in spans) - min(b - a for a, b in spans) > 1:
            fails.append(f"decompose {X}: unbalanced {spans}")
        if spans != [L.decompose_rows(X, world, r) for r in range(world)]:
            fails.append(f"decompose {X}: ranks disagree")

    # 3. per-slab synthetic states of the bench workloads, gathered, equal the monolithic arrays bit for bit
    R, C = 24 * world + 3, 40
    x0, x1 = L.decompose_rows(R, world, rank)
    for name, fn, mono in (("rt", lambda a, b: bench.rt_densities(R, C, a, b), bench.rt_densities(R, C, 0, R)),
                           ("droplet", lambda a, b: bench.droplet_densities(R, R / 4.0, a, b), bench.droplet_densities(R, R / 4.0, 0, R))):
        part = fn(x0, x1)
        for k in range(2):
            got = [None] * world
            dist.all_gather_object(got, part[k])
            if not np.array_equal(np.concatenate(got, axis=0), mono[k]):
                fails.append(f"slab init {name}[{k}]")

    # 4. no CPU fallback on any rank
    if not torch.cuda.is_available():
        try:
            L.Domain(L.default_config(model=L.MODEL_BGK, X=R, Y=C, x0=x0, x1=x1))
            fails.append("lbm_create succeeded without a CUDA device")
        except L.LbmError as e:
            if e.status != 2:  # LBM_ERR_CUDA
                fails.append(f"lbm_create: status {e.status}")

    n = torch.tensor([len(fails)])
    dist.all_reduce(n)
    if fails:
        print(f"rank {rank}: {fails}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(n.item()) else 0)


if __name__ == "__main__":
    main()
