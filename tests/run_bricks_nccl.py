"""Multi-GPU brick check, launched by torchrun (one rank per GPU, NCCL):
every rank deconvolves its brick with halo exchange; rank 0 gathers psi and compares it with the
oracle on the whole volume.  Used by tests/test_multigpu_nccl.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_bricks_nccl.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from spim_registration_b200 import synthetic, bricks
    from oracle import mvdecon_oracle as O
    brick, V, ks, iters = (24, 28, 32), 3, 7, 2
    typ, gen = O.EFFICIENT_BAYESIAN, 2
    grid = bricks.grid_for(world)
    c = bricks.rank_coords(rank, grid)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    _, imgs, ws, psfs = synthetic.make_dataset(gshape, V, ks, kind="beads", seed=3)
    sl = tuple(slice(c[d] * brick[d], (c[d] + 1) * brick[d]) for d in range(3))
    r = bricks.BrickRunner(brick, V, typ, generation=gen, lam=0.006, device=local, rank=rank, world=world,
                           grid=grid, dist=dist)
    for v in range(V):
        r.session.set_view(v, np.ascontiguousarray(imgs[v][sl]), np.ascontiguousarray(ws[v][sl]), psfs[v])
    r.init()
    s, m = r.run(iters, stats=True)
    extra = int(os.environ.get("SPIM_TEST_EXTRA_ITERS", "0"))
    if extra:
        r.run(extra)          # the statistics-free path: one CUDA-graph launch per iteration, exchanges captured
    iters += extra
    r.finish()
    if rank == 0:
        print("EXCHANGE_PATH " + ("p2p" if r.use_p2p else ("pack" if r.use_pack else "slab")))
    psi = torch.from_numpy(r.get_psi()).cuda()
    parts = [torch.empty_like(psi) for _ in range(world)] if rank == 0 else None
    dist.gather(psi, parts, dst=0)
    ok = True
    if rank == 0:
        full = np.zeros(gshape, np.float32)
        for q in range(world):
            cq = bricks.rank_coords(q, grid)
            sq = tuple(slice(cq[d] * brick[d], (cq[d] + 1) * brick[d]) for d in range(3))
            full[sq] = parts[q].cpu().numpy()
        ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=iters, lam=0.006, gen=gen))
        per, l2 = O.parity_errors(full, ref.psi)
        print(f"bricks world={world} grid={grid}: per-voxel {per:.3e} L2 {l2:.3e}")
        ok = per <= 1e-3 and l2 <= 1e-4
        for (it, v, rs, rm) in ref.stats[:s.shape[0] * V]:
            ok = ok and np.isclose(s[it, v], rs, rtol=2e-3, atol=1e-6) and np.isclose(m[it, v], rm, rtol=5e-3, atol=1e-6)
        print("BRICKS_NCCL_OK" if ok else "BRICKS_NCCL_FAIL")
    r.close()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
