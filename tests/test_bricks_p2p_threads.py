"""Direct halo push over peer memory (mvd_p2p_*, SPIM_BRICK_P2P=1) on CPU: world_size 2 / 4 / 8 ranks as THREADS of one
process, each driving the kernel emulator on its brick, so that the ranks really reach each other's buffers through raw
pointers and really run concurrently (ctypes releases the GIL; the emulated wait kernel spins on the flag words).  The package's
in-process group (spim_registration_b200/inprocess.py) supplies the few collectives the runner needs.  The assembled psi must equal the
oracle's result on the whole volume, and every rank must actually have adopted the push path (no silent fallback)."""
import os
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_thread(rank, world, shared, brick, V, ks, typ, gen, iters, lib, data, out, errors):  # noqa: PLR0913
    try:
        from spim_registration_b200 import bricks
        imgs, ws, psfs = data
        grid = bricks.grid_for(world)
        c = bricks.rank_coords(rank, grid)
        sl = tuple(slice(c[d] * brick[d], (c[d] + 1) * brick[d]) for d in range(3))
        r = bricks.BrickRunner(brick, V, typ, generation=gen, lam=0.006, rank=rank, world=world, grid=grid,
                               dist=shared.rank(rank), lib=lib, cpu=True)
        for v in range(V):
            r.session.set_view(v, np.ascontiguousarray(imgs[v][sl]), np.ascontiguousarray(ws[v][sl]), psfs[v])
        r.init()
        used = r.use_p2p
        s, m = r.run(iters, stats=True)
        r.run(1)                       # the statistics-free path (what the benchmark drives) on top
        r.finish()
        out[rank] = dict(psi=r.get_psi(), sl=sl, s=s, m=m, avg=r.session.info().avg, used=used,
                         timed_out=r.session.p2p_timed_out())
        r.close()
    except BaseException as e:         # noqa: BLE001
        errors.append((rank, repr(e)))
        shared.abort()


@pytest.mark.parametrize("world,gen,typ,ks,brick", [(2, 2, 2, 5, (8, 9, 10)), (4, 2, 0, 5, (8, 9, 10)), (8, 1, 1, 5, (8, 9, 10)),
                                                      (8, 2, 2, 7, (8, 9, 10)), (4, 1, 3, 3, (8, 9, 10)),
                                                      # x origin 4, row length 20, x extent 12: the float4 path of the y / z faces
                                                      (8, 2, 2, 7, (8, 8, 12)), (4, 2, 3, 7, (7, 9, 12))])
@pytest.mark.parametrize("fuse", ["1", "0"])
def test_direct_push_bricks_match_whole_volume_oracle(monkeypatch, world, gen, typ, ks, brick, fuse):
    """fuse = 1: the x-inverse epilogue stores the neighbours' halo voxels itself and the push only raises the flags
    (HaloFuse); fuse = 0: the separate copy + signal kernel."""
    import torch
    import __graft_entry__ as g
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import bricks, native, synthetic
    monkeypatch.setenv("SPIM_BRICK_P2P", "1")
    monkeypatch.setenv("SPIM_BRICK_FUSE", fuse)
    torch.set_num_threads(1)
    lib = native.load_library(g.build_emulator())
    fused_before = lib.mvd_debug_counter(3)
    V, iters = 2, 2
    grid = bricks.grid_for(world)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    _, imgs, ws, psfs = synthetic.make_dataset(gshape, V, ks, kind="beads", seed=3)
    from spim_registration_b200.inprocess import ThreadGroup
    shared = ThreadGroup(world)
    out, errors = {}, []
    threads = [threading.Thread(target=_rank_thread, args=(r, world, shared, brick, V, ks, typ, gen, iters, lib,
                                                           (imgs, ws, psfs), out, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    assert len(out) == world
    assert all(o["used"] for o in out.values()), "a rank fell back from the direct-push path"
    fused = lib.mvd_debug_counter(3) - fused_before
    if fuse == "1" and brick[2] % 2 == 0:
        assert fused >= world * V * 2 * (iters + 1), "the fused halo push was not used"
    if fuse == "0":
        assert fused == 0
    assert not any(o["timed_out"] for o in out.values())
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=iters + 1, lam=0.006, gen=gen))
    psi = np.zeros(gshape, np.float32)
    for r in range(world):
        psi[out[r]["sl"]] = out[r]["psi"]
        assert np.isclose(float(out[r]["avg"]), ref.avg, rtol=1e-6)
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)
    for (it, v, rs, rm) in ref.stats:
        if it < iters:
            assert np.isclose(out[0]["s"][it, v], rs, rtol=2e-3, atol=1e-6)
            assert np.isclose(out[0]["m"][it, v], rm, rtol=5e-3, atol=1e-6)


def test_push_api_misuse_is_reported():
    """no connection, wrong order, bad arguments: error codes with messages, never a crash"""
    import __graft_entry__ as g
    from spim_registration_b200 import native, synthetic
    from spim_registration_b200.deconvolution import Session
    lib = native.load_library(g.build_emulator())
    _, imgs, ws, psfs = synthetic.make_dataset((8, 8, 8), 1, 3, kind="beads", seed=1)
    with Session((8, 8, 8), 1, 2, haloed=True, lib=lib) as s:
        with pytest.raises(native.NativeError):
            s.p2p_export()                       # before init
        s.set_view(0, imgs[0], ws[0], psfs[0])
        s.init()
        with pytest.raises(native.NativeError):
            s.p2p_push(0)                        # not connected
        rec = s.p2p_export()
        assert len(rec) == s.P2P_RECORD_BYTES and rec[:8] == b"SPIMP2P1"
        with pytest.raises(native.NativeError):
            s.p2p_connect([bytes(288)], [[0] * 9], [(0, 0)])          # not a record
        with pytest.raises(native.NativeError):
            s.p2p_connect([rec], [[0, 0, 0, 99, 1, 1, 0, 0, 0]], [(0, 0)])   # box out of range
        # a rank may push into itself: slot 5 raised "there" is awaited here
        ptr, dims, origin = s.device_buffer(0)
        s.p2p_connect([rec], [[origin[0], origin[1], origin[2], 1, 1, 1, 0, 0, 0]], [(5, 5)])
        with pytest.raises(native.NativeError):
            s.p2p_wait(0)                        # no outstanding push
        s.p2p_push(0)
        with pytest.raises(native.NativeError):
            s.p2p_push(0)                        # previous push not waited for
        s.p2p_wait(0)
        s.sync()
        assert not s.p2p_timed_out()
        s.p2p_disconnect()
    with Session((8, 8, 8), 1, 2, lib=lib) as s2:
        s2.set_view(0, imgs[0], ws[0], psfs[0])
        s2.init()
        with pytest.raises(native.NativeError):
            s2.p2p_export()                      # not a brick-mode session
