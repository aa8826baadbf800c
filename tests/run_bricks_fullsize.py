"""Full-size multi-GPU brick runs with oracle checks, launched by torchrun (one rank per GPU):

  --config c3   BASELINE configs[2] geometry: bricks of 512x512x256 (8 ranks = 1024x1024x512), 31^3 PSFs, Optimization II.
                ONE full iteration over --views views (default 3), then sub-bricks of psi that straddle brick faces, the
                brick corner in the middle of the volume and a volume corner are recomputed by the oracle from a crop of
                the global inputs.
  --config c5   BASELINE configs[4]: 8 views, 2048x2048x1024 over 8 ranks (bricks of 1024x1024x512), Efficient-Bayesian,
                views handed over cell by cell with mvd_upload_region from device memory.  The first --check-steps
                view-steps (default 2) are verified against the oracle on the same kind of straddling sub-bricks (the
                dependency cone of a whole 8-view iteration would need a 512^3 oracle run), then psi is reset and --iters
                iterations (default 10) are timed.

Set-up and timing are spim_registration_b200/bigvolume.py (shared with bench.py); this file adds the oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      tests/run_bricks_fullsize.py --config c5 --json gpurun_out/c5.json
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


def regions_for(gshape, grid, brick, n):
    """Sub-bricks (name, lo) to verify: the brick corner nearest the volume centre (straddles up to 8 bricks), one face
    centre per split axis (straddles 2), and the volume corner (boundary rule, no neighbour)."""
    out = []
    c = [brick[d] * (grid[d] // 2) if grid[d] > 1 else gshape[d] // 2 for d in range(3)]
    out.append(("corner_of_bricks", [max(0, min(gshape[d] - n, c[d] - n // 2)) for d in range(3)]))
    for d in range(3):
        if grid[d] > 1:
            lo = [max(0, min(gshape[e] - n, (brick[e] // 2 if e != d else c[d] - n // 2))) for e in range(3)]
            out.append((f"face_axis{d}", lo))
    out.append(("volume_corner", [0, 0, 0]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3", choices=["c3", "c5", "custom"])
    ap.add_argument("--brick", type=int, nargs=3, default=None, help="per-rank brick (z y x)")
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--iter-type", type=int, default=None)
    ap.add_argument("--psf", type=int, default=31)
    ap.add_argument("--check-steps", type=int, default=None, help="view-steps verified against the oracle (0 = none)")
    ap.add_argument("--iters", type=int, default=None, help="timed iterations after the check")
    ap.add_argument("--json", default=None)
    ap.add_argument("--region", type=int, default=32)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from spim_registration_b200 import bigvolume, bricks, synthetic
    from spim_registration_b200.deconvolution import PSFTYPE

    if args.config == "c3":
        brick, V, typ, check_steps, iters = (256, 512, 512), 3, int(PSFTYPE.OPTIMIZATION_II), None, 0
    elif args.config == "c5":
        brick, V, typ, check_steps, iters = (512, 1024, 1024), 8, int(PSFTYPE.EFFICIENT_BAYESIAN), 2, 10
    else:
        brick, V, typ, check_steps, iters = (64, 64, 64), 3, int(PSFTYPE.EFFICIENT_BAYESIAN), None, 0
    brick = tuple(args.brick) if args.brick else brick
    V = args.views or V
    typ = args.iter_type if args.iter_type is not None else typ
    check_steps = args.check_steps if args.check_steps is not None else check_steps
    iters = args.iters if args.iters is not None else iters
    if check_steps is None:
        check_steps = V                     # one full iteration
    ks = args.psf

    r, meta = bigvolume.setup_runner(brick, V, typ, rank, world, local, dist if world > 1 else None, psf_size=ks)
    s = r.session
    psfs = meta.pop("psfs")
    gshape, grid = tuple(meta["global_zyx"]), tuple(meta["grid_zyx"])
    result = dict(meta, config=args.config)

    # ---------------- oracle check of the first view-steps on straddling sub-bricks ----------------------
    ok = True
    if check_steps > 0:
        n = args.region
        for v in range(check_steps):          # view-steps in view order, exactly BrickRunner._iteration
            if world > 1:
                r.exchange(0)
            s.view_phase(v, 0)
            if world > 1:
                r.exchange(1)
            s.view_phase(v, 1)
        if check_steps == V:
            r.finish()
        s.sync()
        checks = []
        for name, lo in regions_for(gshape, grid, brick, n):
            box = bigvolume.gather_region(r, meta, lo, n, dist if world > 1 else None)
            if rank == 0:
                from oracle import mvdecon_oracle as O
                h = (ks // 2) * 2 * check_steps          # every view-step reaches 2 * (k // 2) voxels further
                clo = [max(0, lo[d] - h) for d in range(3)]
                chi = [min(gshape[d], lo[d] + n + h) for d in range(3)]
                ext = [chi[d] - clo[d] for d in range(3)]
                k1, k2 = O.init_kernels(psfs, typ)
                p = O.DeconParams(iteration_type=typ, lam=0.006, gen=O.GEN2)
                psi = np.full(ext, np.float32(meta["avg"]), np.float32)
                for v in range(check_steps):
                    im, w = synthetic.hash_view(gshape, clo, ext, v, V, xp="numpy")
                    psi, _, _ = O.view_step(psi, im, w, k1[v], k2[v], p)
                rs = tuple(slice(lo[d] - clo[d], lo[d] - clo[d] + n) for d in range(3))
                per, l2 = O.parity_errors(box, psi[rs])
                good = bool(per <= 1e-3 and l2 <= 1e-4)
                ok = ok and good
                checks.append({"region": name, "lo_zyx": [int(x) for x in lo], "size": n, "per_voxel": float(per),
                               "rel_l2": float(l2), "ok": good})
                print(f"[fullsize {args.config}] {name} at {lo}: per-voxel {per:.3e} L2 {l2:.3e} {'ok' if good else 'FAIL'}", flush=True)
        result["oracle_check"] = {"view_steps": check_steps, "regions": checks,
                                  "tolerance": "per-voxel <= 1e-3, relative L2 <= 1e-4 (BASELINE north_star)"}
        if iters > 0:                           # back to the initial estimate for the timed part
            s.set_avg(meta["avg"], float(getattr(r, "osem", 1.0)))
            r._p2p_last = None
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()

    # ---------------- timed iterations --------------------------------------------------------------------
    if iters > 0:
        ms = bigvolume.time_iterations(r, iters, dist if world > 1 else None)
        result.update({"iterations": iters, "ms_per_iteration": ms,
                       "value": int(np.prod(gshape)) * V / (ms * 1e-3), "unit": "voxel-view-iters/s",
                       "timing": "CUDA events on the session stream, max over ranks, one warm-up iteration before"})
    result["peak_device_bytes_per_gpu"] = bigvolume.peak_device_bytes(dist if world > 1 else None, world)
    result["hbm_total_bytes"] = int(torch.cuda.mem_get_info()[1])
    result["ok"] = bool(ok)
    if rank == 0:
        line = json.dumps(result)
        print(line, flush=True)
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)) or ".", exist_ok=True)
            with open(args.json, "w") as f:
                f.write(line + "\n")
        print("FULLSIZE_OK" if ok else "FULLSIZE_FAIL", flush=True)
    r.close()
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
