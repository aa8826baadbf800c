"""Multi-process brick partition + halo exchange on CPU: world_size 2 / 4 / 8 over gloo, each rank
driving the kernel emulator on its brick.  The assembled psi must equal the oracle's result on the
whole (unpartitioned) volume -- the reference's 'blocked == unblocked' property, distributed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, brick, V, ks, typ, gen, iters, emu_path, outdir, pack="1"):
    sys.path.insert(0, ROOT)
    os.environ["SPIM_BRICK_PACK"] = pack
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spim_registration_b200 import native, synthetic, bricks
    lib = native.load_library(emu_path)
    grid = bricks.grid_for(world)
    c = bricks.rank_coords(rank, grid)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    _, imgs, ws, psfs = synthetic.make_dataset(gshape, V, ks, kind="beads", seed=3)
    sl = tuple(slice(c[d] * brick[d], (c[d] + 1) * brick[d]) for d in range(3))
    r = bricks.BrickRunner(brick, V, typ, generation=gen, lam=0.006, rank=rank, world=world, grid=grid,
                           dist=dist, lib=lib, cpu=True)
    for v in range(V):
        r.session.set_view(v, np.ascontiguousarray(imgs[v][sl]), np.ascontiguousarray(ws[v][sl]), psfs[v])
    r.init()
    s, m = r.run(iters, stats=True)
    r.finish()
    psi = r.get_psi()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), psi=psi, sl=np.array([[x.start, x.stop] for x in sl]), s=s, m=m,
             avg=r.session.info().avg)
    r.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,gen,typ,pack,ks", [(2, 2, 2, "1", 5), (4, 2, 0, "1", 5), (8, 1, 1, "1", 5), (2, 2, 2, "0", 5),
                                                   (4, 1, 3, "0", 5), (4, 2, 2, "1", 7), (2, 2, 3, "0", 3)])
def test_bricks_match_whole_volume_oracle(tmp_path, world, gen, typ, pack, ks):
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import synthetic, bricks
    emu = g.build_emulator()
    # ks = 7 / 3: odd PSF/2 halos (3 / 1), where the haloed buffers pad their x origin to an even index
    brick, V, iters = (8, 9, 10), 2, 2
    port = _free_port()
    # pack = "0": the slab-copy exchange (the fallback of the single-launch pack / unpack path) must be just as correct
    mp.spawn(_worker, args=(world, port, brick, V, ks, typ, gen, iters, emu, str(tmp_path), pack), nprocs=world, join=True)
    grid = bricks.grid_for(world)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    _, imgs, ws, psfs = synthetic.make_dataset(gshape, V, ks, kind="beads", seed=3)
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=iters, lam=0.006, gen=gen))
    psi = np.zeros(gshape, np.float32)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        sl = tuple(slice(a, b) for a, b in d["sl"])
        psi[sl] = d["psi"]
        assert np.isclose(float(d["avg"]), ref.avg, rtol=1e-6)
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)
    d0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    for (it, v, rs, rm) in ref.stats:
        assert np.isclose(d0["s"][it, v], rs, rtol=2e-3, atol=1e-6)
        assert np.isclose(d0["m"][it, v], rm, rtol=5e-3, atol=1e-6)


def test_grid_and_rank_mapping():
    from spim_registration_b200 import bricks
    assert bricks.grid_for(1) == (1, 1, 1) and bricks.grid_for(2) == (1, 1, 2)
    assert bricks.grid_for(4) == (1, 2, 2) and bricks.grid_for(8) == (2, 2, 2)
    for w in (1, 2, 3, 4, 6, 8, 12):
        g = bricks.grid_for(w)
        assert g[0] * g[1] * g[2] == w
        seen = set()
        for r in range(w):
            c = bricks.rank_coords(r, g)
            assert bricks.coords_rank(c, g) == r
            seen.add(c)
        assert len(seen) == w


def _fusion_worker(rank, world, port, brick, V, emu_path, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spim_registration_b200 import native, synthetic, bricks
    import fusion_cases as FC
    lib = native.load_library(emu_path)
    grid = bricks.grid_for(world)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    stacks, models = FC.make_view_set(V, (8, 18, 22), gshape, seed=5)
    psfs = synthetic.make_psfs(V, 3)
    r = bricks.BrickRunner(brick, V, 3, generation=2, lam=0.006, rank=rank, world=world, grid=grid, dist=dist, lib=lib, cpu=True)
    mn, avg = r.fuse_stacks(stacks, models, (-1, 0, 1), (1, 1, 0), (4, 4, 2), virtual=True, psfs=psfs)
    r.init()
    r.run(2, stats=True)
    r.finish()
    c = bricks.rank_coords(rank, grid)
    sl = tuple(slice(c[d] * brick[d], (c[d] + 1) * brick[d]) for d in range(3))
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), psi=r.get_psi(), sl=np.array([[x.start, x.stop] for x in sl]), mn=mn, avg=avg)
    r.close()
    dist.barrier()
    dist.destroy_process_group()


def test_bricks_fuse_raw_stacks_then_deconvolve(tmp_path):
    """Raw stacks -> per-brick device-side transformation + weights -> brick deconvolution (world 4) equals the
    single-volume pipeline of the oracle."""
    import torch.multiprocessing as mp
    import __graft_entry__ as g
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fusion_cases as FC
    from oracle import fusion_oracle as F
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import synthetic, bricks
    emu = g.build_emulator()
    world, brick, V = 4, (12, 8, 10), 2
    mp.spawn(_fusion_worker, args=(world, _free_port(), brick, V, emu, str(tmp_path)), nprocs=world, join=True)
    grid = bricks.grid_for(world)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    stacks, models = FC.make_view_set(V, (8, 18, 22), gshape, seed=5)
    imgs, ws = FC.oracle_views(stacks, models, gshape, (-1, 0, 1), (1, 1, 0), (4, 4, 2))
    sumw, mn, avg = F.weight_normalizer_virtual(ws, 1)
    wv = [F.normalizing_access(w, sumw, 1.0) for w in ws]
    psfs = synthetic.make_psfs(V, 3)
    ref = O.deconvolve(imgs, wv, psfs, O.DeconParams(iteration_type=3, num_iterations=2, lam=0.006, gen=2))
    psi = np.zeros(gshape, np.float32)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        psi[tuple(slice(a, b) for a, b in d["sl"])] = d["psi"]
        assert int(d["mn"]) == mn and abs(float(d["avg"]) - avg) < 1e-12
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)
