"""Multi-GPU brick partition with halo exchange -- one process per GPU.

The reference already proves locality: a block's result depends only on the block plus a
kernel-sized overlap (mpicbg/spim/postprocessing/deconvolution2/Block.java:385-398,437-447;
spim/process/cuda/BlockGeneratorFixedSizePrecise.java:46-122), and its multi-device mode hands
blocks to one Java thread per device with all data returning to host RAM after every block
(MVDeconFFT.java:447-469).  Here the volume is cut once into persistent bricks (2x2x2 for 8 GPUs,
2x2x1 for 4, 2x1x1 for 2 over (x,y,z)); each rank keeps its brick of psi, of every view image and
weight, and the kernel spectra on its own GPU for the whole run.  Before conv1 the psi halo
(width PSF/2) and before conv2 the ratio halo are refreshed: neighbour faces travel over NCCL
(torch.distributed batch_isend_irecv), volume faces are filled by the convolution's own
out-of-bounds rule (mvd_fill_halo).  Axes are processed x, then y, then z with full extents, so
edges and corners ride along.  No other collective is on the data path; the change statistics
and the initial average are 2-6 scalar all-reduces.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from .deconvolution import Session


def grid_for(world: int) -> Tuple[int, int, int]:
    """Bricks along (z, y, x).  x and y are split first: the per-GPU bench brick is 512x512x256, so
    this keeps the global volume's aspect (8 -> 1024x1024x512)."""
    return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world) or _factor(world)


def _factor(world: int) -> Tuple[int, int, int]:
    g = [1, 1, 1]
    ax = 2
    w = world
    p = 2
    while w > 1:
        while w % p:
            p += 1
        g[ax] *= p
        w //= p
        ax = (ax - 1) % 3
    return tuple(g)


def rank_coords(rank: int, grid: Sequence[int]) -> Tuple[int, int, int]:
    """rank -> (cz, cy, cx), x fastest."""
    cx = rank % grid[2]
    cy = (rank // grid[2]) % grid[1]
    cz = rank // (grid[2] * grid[1])
    return (cz, cy, cx)


def coords_rank(c: Sequence[int], grid: Sequence[int]) -> int:
    return (c[0] * grid[1] + c[1]) * grid[2] + c[2]


class BrickRunner:
    """Drives one brick.  world == 1 degenerates to the plain device-resident session (mvd_run)."""

    def __init__(self, brick: Sequence[int], num_views: int, iteration_type: int, generation: int = 2,
                 lam: float = 0.006, osem_speedup: float = 1.0, device: int = 0, rank: int = 0, world: int = 1,
                 grid: Optional[Sequence[int]] = None, dist=None, lib=None, cpu: bool = False, osem_index: int = 0):
        self.world = world
        self.rank = rank
        self.grid = tuple(grid) if grid else grid_for(world)
        self.coords = rank_coords(rank, self.grid)
        self.dist = dist
        self.cpu = cpu
        self.device = device
        self.num_views = num_views
        self.generation = generation
        self.osem_speedup = osem_speedup
        self.osem_index = osem_index
        self.haloed = world > 1
        self.session = Session(brick, num_views, iteration_type, generation=generation, lam=lam,
                               osem_speedup=osem_speedup, device=device, haloed=self.haloed, lib=lib)
        self._bufs = None
        self.use_pack = False
        self.use_p2p = False
        self._tmp = {}
        self._graph = None
        self._graph_failed = False
        self.use_graph = (not cpu) and os.environ.get("SPIM_BRICK_GRAPH", "1") != "0"

    # ------------------------------------------------------------------------------------------
    def _wrap(self, ptr: int, dims):
        import torch
        if self.cpu:
            n = int(np.prod(dims))
            arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n,)).reshape(dims)
            return torch.from_numpy(arr)

        class _CAI:
            pass
        o = _CAI()
        o.__cuda_array_interface__ = {"shape": tuple(dims), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(o, device=torch.device("cuda", self.device))

    def _stream_ctx(self):
        import contextlib
        if self.cpu:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(torch.cuda.ExternalStream(self.session.stream(), device=torch.device("cuda", self.device)))

    def init(self):
        s = self.session
        s.init()
        if not self.haloed:
            return
        import torch
        part = s.init_partials()
        if self.dist is not None:
            dev = "cpu" if self.cpu else torch.device("cuda", self.device)
            t = torch.tensor(part[:4], dtype=torch.float64, device=dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            m = torch.tensor([part[4]], dtype=torch.float64, device=dev)
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MIN)
            part = np.concatenate([t.cpu().numpy(), m.cpu().numpy(), [0.0]])
        if self.generation == 2:
            avg = part[0] / part[1] if part[1] > 0 else 0.5
            osem = self.osem_speedup
        else:
            avg = float(np.float32(part[0] / part[1])) if part[1] > 0 else 1.0
            osem = self.osem_speedup
            # BayesMVDeconvolution.java:94-97: osemspeedupindex 1 = min #overlapping views, 2 = avg #overlapping views
            # (both >= 1), from the all-reduced counts exactly like mvd_init does for a whole volume
            if self.osem_index == 1:
                osem = max(1.0, float(int(part[4])))
            elif self.osem_index == 2:
                osem = max(1.0, part[2] / part[3]) if part[3] > 0 else 1.0
        self.avg, self.osem = avg, osem
        s.set_avg(avg, osem)
        bufs = []
        for which in (0, 1):
            ptr, dims, origin = s.device_buffer(which)
            bufs.append((self._wrap(ptr, dims), dims, origin))
        self._bufs = bufs
        i = s.info()
        self.halo_lo = tuple(i.halo_lo)
        self.halo_hi = tuple(i.halo_hi)
        for d in range(3):
            if self.grid[d] > 1:
                # a neighbour's halo is filled from MY interior cells only, and both directions are exchanged as a pair
                if max(self.halo_lo[d], self.halo_hi[d]) > s.dims[d]:
                    raise ValueError(f"brick extent {s.dims[d]} along axis {d} is smaller than the halo "
                                     f"({self.halo_lo[d]}, {self.halo_hi[d]}): use fewer bricks along this axis")
                if (self.halo_lo[d] == 0) != (self.halo_hi[d] == 0):
                    raise NotImplementedError(f"PSF extent 2 along a split axis (halo {self.halo_lo[d]} / {self.halo_hi[d]}): "
                                              "one-sided halos are not exchanged; use an odd PSF size or do not split this axis")
        self._plan_exchange()

    # ------------------------------------------------------------------------------------------
    def fuse_stacks(self, stacks, transforms, bb_min, blending_border, blending_range, virtual: bool = True,
                    psfs=None, normalize_stacks: bool = True):
        """Device-side fusion pre-step for this rank's brick (spim_fusion.h; ProcessForDeconvolution.java:180-366):
        every rank loads the raw stacks and resamples them into ITS brick of the bounding box -- the brick origin is
        simply added to the bounding-box offset, the per-voxel arithmetic is position-independent -- then the weights are
        normalised voxel-locally.  Returns the global (min, avg) number of overlapping views: counts are all-reduced, so
        avg is the reference's value for a single portion (Threads.numThreads() * 2 == 1)."""
        from . import fusion
        s = self.session
        n = s.dims
        origin_xyz = [self.coords[2] * n[2], self.coords[1] * n[1], self.coords[0] * n[0]]
        off = [int(bb_min[d]) + origin_xyz[d] for d in range(3)]
        for v in range(self.num_views):
            st = np.ascontiguousarray(stacks[v], dtype=np.float32)
            fusion.load_stack(s, st, normalize=normalize_stacks)
            bl = fusion.Blending(st.shape[::-1], blending_border, blending_range)
            fusion.transform_view(s, v, transforms[v], off, bl)
            if psfs is not None:
                fusion.set_psf(s, v, psfs[v])
        fusion.load_stack(s, None)
        mn, avg = fusion.normalize_weights(s, virtual=virtual, num_portions=1)
        if self.dist is not None and self.world > 1:
            import torch
            dev = "cpu" if self.cpu else torch.device("cuda", self.device)
            nvox = float(np.prod(n))
            t = torch.tensor([avg * nvox, nvox], dtype=torch.float64, device=dev)      # avg * nvox = exact integer count
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            m = torch.tensor([float(mn)], dtype=torch.float64, device=dev)
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MIN)
            mn, avg = int(m.item()), float(t[0].item() / t[1].item())
        return mn, avg

    def _neighbour(self, axis: int, step: int) -> Optional[int]:
        c = list(self.coords)
        c[axis] += step
        if c[axis] < 0 or c[axis] >= self.grid[axis]:
            return None
        return coords_rank(c, self.grid)

    def _boxes(self, off):
        """(send, recv) slices of my haloed buffers for the neighbour at offset ``off`` in {-1,0,1}^3 (z, y, x), or None when
        a halo involved has zero width.  Every brick has the same geometry, so the same formula describes any rank."""
        n = self.session.dims
        _, dims, origin = self._bufs[0]
        send, recv = [], []
        for d in range(3):
            o, wlo, whi = origin[d], self.halo_lo[d], self.halo_hi[d]
            if d == 2 and off[d] != 0 and wlo > 0 and whi > 0 and os.environ.get("SPIM_BRICK_XPAD", "1") != "0":
                # x pieces widened to whole 16-byte groups where the buffer has the cells (a 15-voxel halo moves as 16 columns):
                # the push kernel then moves them as float4 like the y / z faces; the extra column lands in a cell nobody reads
                w4lo, w4hi = (wlo + 3) & ~3, (whi + 3) & ~3
                if o % 4 == 0 and n[d] % 4 == 0 and o >= w4lo and o + n[d] + w4hi <= dims[d] and n[d] >= max(w4lo, w4hi):
                    wlo, whi = w4lo, w4hi
            if off[d] == 0:
                send.append(slice(o, o + n[d])); recv.append(slice(o, o + n[d]))
            elif off[d] < 0:       # neighbour below: it needs my first `whi` interior cells; I get its last `wlo`
                if whi == 0 or wlo == 0:
                    return None
                send.append(slice(o, o + whi)); recv.append(slice(o - wlo, o))
            else:                   # neighbour above: it needs my last `wlo` interior cells; I get its first `whi`
                if whi == 0 or wlo == 0:
                    return None
                send.append(slice(o + n[d] - wlo, o + n[d])); recv.append(slice(o + n[d], o + n[d] + whi))
        return tuple(send), tuple(recv)

    def _plan_exchange(self):
        """Neighbour list for the one-shot exchange: every existing neighbour at offset (dz,dy,dx) in
        {-1,0,1}^3 gets the interior cells it needs (face, edge or corner piece) and sends back the
        matching piece for my halo.  Volume faces need nothing: the convolution loader applies the
        out-of-bounds rule there (mvd_set_halo_mask)."""
        import itertools
        plan, offs = [], []
        for off in itertools.product((-1, 0, 1), repeat=3):
            if off == (0, 0, 0):
                continue
            c = [self.coords[d] + off[d] for d in range(3)]
            if any(c[d] < 0 or c[d] >= self.grid[d] for d in range(3)):
                continue
            b = self._boxes(off)
            if b is not None:
                plan.append((coords_rank(c, self.grid), b[0], b[1]))
                offs.append(off)
        self._xoffs = offs
        self._xplan = plan
        # flat staging buffers for the single-launch pack / unpack path
        import torch
        t = self._bufs[0][0]
        box = lambda sl_: tuple(s_.start for s_ in sl_) + tuple(s_.stop - s_.start for s_ in sl_)
        self._send_regions = [box(ssl) for _, ssl, _ in plan]
        self._recv_regions = [box(rsl) for _, _, rsl in plan]
        sizes_s = [int(np.prod(r[3:])) for r in self._send_regions]
        sizes_r = [int(np.prod(r[3:])) for r in self._recv_regions]
        self._send_flat = torch.empty(max(1, sum(sizes_s)), dtype=t.dtype, device=t.device)
        self._recv_flat = torch.empty(max(1, sum(sizes_r)), dtype=t.dtype, device=t.device)
        self._send_off = np.concatenate([[0], np.cumsum(sizes_s)]).astype(int)
        self._recv_off = np.concatenate([[0], np.cumsum(sizes_r)]).astype(int)
        # sides with a neighbour read the exchanged halo; all other sides (volume faces) use the out-of-bounds rule
        lo_mask = sum(1 << d for d in range(3) if self._neighbour(d, -1) is not None)
        hi_mask = sum(1 << d for d in range(3) if self._neighbour(d, +1) is not None)
        self.session.set_halo_mask(lo_mask, hi_mask)
        self.use_pack = os.environ.get("SPIM_BRICK_PACK", "1") != "0"
        if self.use_pack and plan and self.dist is not None:
            self.use_pack = self._verify_mode("pack")
            if not self.use_pack and self.rank == 0:
                print("[bricks] halo exchange: using slab copies (pack-path self-check failed)", flush=True)
        # direct halo push over NVLink peer memory (mvd_p2p_*): the default; the NCCL batch is the fallback when a peer cannot be
        # mapped or the one-time self-check fails (SPIM_BRICK_P2P=0 forces the NCCL path)
        self.use_p2p = False
        self._p2p_last = None
        if os.environ.get("SPIM_BRICK_P2P", "1") == "1" and plan and self.dist is not None:
            self._setup_p2p()

    def _setup_p2p(self):
        """Exchange the export records, connect every neighbour piece, and adopt the push path only if EVERY rank connected
        and the pushed halos equal the slab-copy exchange bit for bit (otherwise all ranks stay on the NCCL path)."""
        import torch
        s = self.session
        dev = "cpu" if self.cpu else torch.device("cuda", self.device)
        ok = True
        try:
            rec = s.p2p_export()
        except Exception as e:                    # noqa: BLE001
            rec, ok = bytes(s.P2P_RECORD_BYTES), False
            if self.rank == 0:
                print(f"[bricks] direct halo push unavailable ({type(e).__name__}: {e})", flush=True)
        mine = torch.tensor(list(rec), dtype=torch.uint8, device=dev)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(every, mine)
        if ok:
            try:
                slot = lambda o: (o[0] + 1) * 9 + (o[1] + 1) * 3 + (o[2] + 1)
                records, boxes, slots = [], [], []
                for (peer, ssl, _), off in zip(self._xplan, self._xoffs):
                    neg = tuple(-o for o in off)
                    dst = self._boxes(neg)[1]                  # where the neighbour receives what comes from my direction
                    records.append(bytes(every[peer].cpu().numpy().tobytes()))
                    boxes.append([sl.start for sl in ssl] + [sl.stop - sl.start for sl in ssl] + [sl.start for sl in dst])
                    slots.append((slot(neg), slot(off)))
                s.p2p_connect(records, boxes, slots)
            except Exception as e:                # noqa: BLE001
                ok = False
                if self.rank == 0:
                    print(f"[bricks] direct halo push unavailable ({type(e).__name__}: {e})", flush=True)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if not int(flag.item()):
            return
        self.use_p2p = self._verify_mode("p2p")
        if self.rank == 0:
            print("[bricks] halo exchange: direct push over peer memory" if self.use_p2p else
                  "[bricks] halo exchange: direct-push self-check failed, staying on the NCCL path", flush=True)

    def _verify_mode(self, mode: str) -> bool:
        """One-time self-check of an exchange path ("pack": single-launch pack / unpack around one NCCL batch; "p2p": direct
        push over peer memory): it must reproduce the halo the plain slab-copy exchange produces, bit for bit, on every
        rank -- otherwise all ranks fall back together."""
        import torch
        t = self._bufs[1][0]                      # the ratio buffer is scratch at this point
        keep = t.clone()
        if self.cpu:
            g = torch.Generator().manual_seed(1234 + self.rank)
            fill = torch.rand(t.shape, generator=g, dtype=t.dtype)
        else:
            g = torch.Generator(device=t.device).manual_seed(1234 + self.rank)
            fill = torch.rand(t.shape, generator=g, device=t.device, dtype=t.dtype)
        t.copy_(fill)
        saved = (self.use_pack, self.use_p2p)
        ok = True
        try:
            # torch fills the buffer on ITS current stream; the exchange runs on the session's (non-blocking) stream:
            # synchronise the device between the two, or the exchange may pack the buffer before it is filled
            self._sync_stream()
            self.use_pack, self.use_p2p = False, False
            self.exchange(1)
            self._sync_stream()
            want = t.clone()
            t.copy_(fill)
            self._sync_stream()
            self.dist.barrier()                   # nobody may push into a halo its owner is still reading
            self.use_pack, self.use_p2p = (mode == "pack"), (mode == "p2p")
            self.exchange(1)
            self._sync_stream()
            ok = bool(torch.equal(t, want))
            if mode == "p2p" and self.session.p2p_timed_out():
                ok = False
        except Exception as e:                    # argument / launch errors surface before any NCCL call is posted
            ok = False
            if self.rank == 0:
                print(f"[bricks] {mode} path unavailable ({type(e).__name__}: {e})", flush=True)
        self.use_pack, self.use_p2p = saved
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=t.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        self.dist.barrier()
        t.copy_(keep)
        self._sync_stream()
        self._p2p_last = None
        return bool(int(flag.item()))

    def _sync_stream(self):
        self.session.sync()
        if not self.cpu:
            import torch
            torch.cuda.synchronize()

    def exchange(self, which: int):
        """Refresh the halo of buffer ``which`` (0 = psi, 1 = ratio) from all neighbours in ONE batch of
        NCCL send/recv (faces, edges and corners together)."""
        import torch
        t, dims, origin = self._bufs[which]
        if not self._xplan:
            return
        if self.use_p2p:
            # one fused copy + signal kernel into the neighbours' halos, one wait kernel; nothing else.  Pushes of the two
            # buffers alternate inside an iteration, which is what orders a push after the neighbour's last read of that
            # halo; anything else (two pushes of one buffer in a row) is separated by a barrier across ranks.
            if self._p2p_last == which:
                self._sync_stream()
                self.dist.barrier()
            self._p2p_last = which
            self.session.p2p_push(which)
            self.session.p2p_wait(which)
            return
        if self.use_pack:
            # one gather kernel, one NCCL batch on slices of the flat buffers, one scatter kernel
            with self._stream_ctx():
                self.session.halo_pack(which, self._send_regions, self._send_flat.data_ptr())
                ops = []
                for i, (peer, _, _) in enumerate(self._xplan):
                    ops.append(self.dist.P2POp(self.dist.isend, self._send_flat[self._send_off[i]:self._send_off[i + 1]], peer))
                    ops.append(self.dist.P2POp(self.dist.irecv, self._recv_flat[self._recv_off[i]:self._recv_off[i + 1]], peer))
                for r in self.dist.batch_isend_irecv(ops):
                    r.wait()
                self.session.halo_pack(which, self._recv_regions, self._recv_flat.data_ptr(), unpack=True)
            return
        with self._stream_ctx():
            ops, recvs = [], []
            for peer, ssl, rsl in self._xplan:
                send = t[ssl].contiguous()
                buf = torch.empty(tuple(s_.stop - s_.start for s_ in rsl), dtype=t.dtype, device=t.device)
                ops.append(self.dist.P2POp(self.dist.isend, send, peer))
                ops.append(self.dist.P2POp(self.dist.irecv, buf, peer))
                recvs.append((buf, rsl))
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()
            for buf, rsl in recvs:
                t[rsl].copy_(buf)

    def _iteration(self, stats: bool = False, out=None):
        for v in range(self.num_views):
            self.exchange(0)
            self.session.view_phase(v, 0)
            self.exchange(1)
            st = self.session.view_phase(v, 1, want_stats=stats)
            if stats:
                out.append(st)

    def _capture(self):
        """Capture one whole iteration (all kernels of every view-step + the NCCL halo exchanges) into a
        CUDA graph on the session's stream: the host then issues ONE launch per iteration."""
        import torch
        try:
            stream = torch.cuda.ExternalStream(self.session.stream(), device=torch.device("cuda", self.device))
            self._iteration()            # warm-up outside capture (allocations, index tables, NCCL channels)
            self.session.sync()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                self._iteration()
            self._graph = g
        except Exception as e:       # stay correct: fall back to eager launches
            self._graph = None
            self._graph_failed = True
            if self.rank == 0:
                print(f"[bricks] CUDA-graph capture unavailable ({type(e).__name__}: {e}); running eagerly", flush=True)
            torch.cuda.synchronize()

    def run(self, n_iterations: int, stats: bool = False):
        if not self.haloed:
            return self.session.run(n_iterations, stats=stats)
        if stats or not self.use_graph:
            out = []
            for _ in range(n_iterations):
                self._iteration(stats, out)
        else:
            done = 0
            if self._graph is None and not self._graph_failed:
                self._capture()          # runs one eager iteration as its warm-up
                done = 1 if n_iterations > 0 else 0
                if n_iterations == 0:
                    raise RuntimeError("BrickRunner.run(0) before the first real iteration is not supported")
            for _ in range(n_iterations - done):
                if self._graph is not None:
                    self._graph.replay()
                else:
                    self._iteration()
            out = None
        if stats:
            import torch
            a = np.array(out, dtype=np.float64).reshape(n_iterations, self.num_views, 2)
            if self.dist is not None:
                dev = "cpu" if self.cpu else torch.device("cuda", self.device)
                s = torch.tensor(a[..., 0], dtype=torch.float64, device=dev)
                m = torch.tensor(a[..., 1], dtype=torch.float64, device=dev)
                self.dist.all_reduce(s, op=self.dist.ReduceOp.SUM)
                self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
                return s.cpu().numpy(), m.cpu().numpy()
            return a[..., 0], a[..., 1]
        self.session.sync()
        if not self.cpu:
            import torch
            torch.cuda.current_stream().synchronize()
        if self.use_p2p and self.session.p2p_timed_out():
            raise RuntimeError("direct halo push: a neighbour's halo did not arrive within SPIM_P2P_TIMEOUT_S seconds")
        return None

    def extra_launches_per_iteration(self) -> int:
        if not self.haloed:
            return 0
        if getattr(self, "use_p2p", False) or getattr(self, "use_pack", False):
            return 4 * self.num_views        # push + wait, or pack + unpack, per exchange; two exchanges per view-step
        return 0

    def finish(self):
        self.session.finish()

    def get_psi(self) -> np.ndarray:
        return self.session.get_psi()

    def close(self):
        # release the captured graph (it references NCCL work) before the session and the process group go away
        if self._graph is not None:
            import torch
            torch.cuda.synchronize()
            self._graph = None
        self._bufs = None
        if getattr(self, "use_p2p", False):
            # nobody unmaps or frees a buffer a neighbour may still be pushing into
            self._sync_stream()
            self.dist.barrier()
            self.session.p2p_disconnect()
            self.dist.barrier()
            self.use_p2p = False
        self.session.close()
