"""Several bricks driven from ONE process -- one host thread per device, like the reference's multi-device mode
(MVDeconFFT.java:447-469: one Java thread per entry of deviceList) -- instead of one process per GPU under torchrun.

`ThreadGroup` supplies the few collectives `bricks.BrickRunner` needs (barrier, all_reduce, all_gather and the slab
send / receive pairs of the fallback exchange) between the threads of a process.  With SPIM_BRICK_P2P=1 the data path then
needs none of them: same-process sessions reach each other's buffers through raw device pointers (the export records carry
the process id; mvd_p2p_connect enables peer access), and the halo exchange is the fused push + wait kernels alone.

    group = ThreadGroup(world)
    threads = [threading.Thread(target=rank_main, args=(group.rank(r), r)) for r in range(world)]
    # rank_main(dist, r): bricks.BrickRunner(..., device=r, rank=r, world=world, dist=dist) ... as under torchrun

The same class backs the CPU tests of the push protocol (tests/test_bricks_p2p_threads.py), where every rank drives the
kernel emulator."""
from __future__ import annotations

import collections
import threading


class _Work:
    def wait(self):
        return True


class _RankView:
    """What one rank sees: the subset of torch.distributed that BrickRunner uses."""

    class ReduceOp:
        SUM, MIN, MAX = "sum", "min", "max"

    isend, irecv = "isend", "irecv"

    def __init__(self, group: "ThreadGroup", rank: int):
        self.g, self.rank_ = group, rank

    @staticmethod
    def P2POp(op, tensor, peer):
        return (op, tensor, peer)

    def barrier(self):
        self.g._barrier.wait()

    def all_reduce(self, t, op=None):
        import torch
        self.g._slots[self.rank_] = t.detach().to("cpu", copy=True)
        self.g._barrier.wait()
        st = torch.stack(list(self.g._slots))
        res = {"sum": st.sum(0), "min": st.min(0).values, "max": st.max(0).values}[op or "sum"]
        self.g._barrier.wait()
        t.copy_(res.to(t.device))

    def all_gather(self, out, t):
        self.g._slots[self.rank_] = t.detach().to("cpu", copy=True)
        self.g._barrier.wait()
        for r in range(self.g.world):
            out[r].copy_(self.g._slots[r].to(out[r].device))
        self.g._barrier.wait()

    def batch_isend_irecv(self, ops):
        """Every rank calls this the same number of times (each brick has at least one neighbour for world > 1)."""
        for op, t, peer in ops:
            if op == "isend":
                with self.g._lock:
                    self.g._mail[(self.rank_, peer)].append(t.detach().to("cpu", copy=True))
        self.g._barrier.wait()
        for op, t, peer in ops:
            if op == "irecv":
                with self.g._lock:
                    t.copy_(self.g._mail[(peer, self.rank_)].popleft().to(t.device))
        self.g._barrier.wait()
        return [_Work() for _ in ops]


class ThreadGroup:
    def __init__(self, world: int):
        self.world = int(world)
        self._barrier = threading.Barrier(self.world)
        self._slots = [None] * self.world
        self._mail = collections.defaultdict(collections.deque)
        self._lock = threading.Lock()

    def rank(self, r: int) -> _RankView:
        return _RankView(self, r)

    def abort(self):
        """Release every rank waiting in a collective (call when one rank failed)."""
        self._barrier.abort()
