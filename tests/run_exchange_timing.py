"""Halo-exchange timing on real links, launched by torchrun (one rank per GPU): bricks of the bench size, the direct push
(mvd_p2p_push + mvd_p2p_wait) of the psi buffer timed alone with CUDA events on the session stream, max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
      tests/run_exchange_timing.py [--brick 256 512 512] [--reps 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--brick", type=int, nargs=3, default=[256, 512, 512])
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from spim_registration_b200 import bigvolume
    r, meta = bigvolume.setup_runner(tuple(args.brick), args.views, 2, rank, world, local, dist)
    s = r.session
    stream = torch.cuda.ExternalStream(s.stream(), device=dev)
    out = {"world": world, "brick_zyx": args.brick, "exchange": meta["exchange"], "xpad": os.environ.get("SPIM_BRICK_XPAD", "1")}
    nbytes = sum(int(np.prod(reg[3:])) * 4 for reg in r._send_regions)
    out["bytes_sent_per_rank"] = nbytes
    out["pieces"] = len(r._send_regions)
    if r.use_p2p:
        # pushes of the two buffers alternate, exactly as inside an iteration (two pushes of one buffer in a row would need a
        # barrier across ranks, bricks.py)
        for _ in range(5):
            r.exchange(0); r.exchange(1)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            r.exchange(0); r.exchange(1)
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / (2 * args.reps)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["us_per_exchange"] = float(t.item()) * 1e3
        out["gbs_sent_per_rank"] = nbytes / (float(t.item()) * 1e-3) / 1e9
    if rank == 0:
        print("EXCHANGE_TIMING " + json.dumps(out), flush=True)
    r.close()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
