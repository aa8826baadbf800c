#!/usr/bin/env python
"""bench.py -- multi-view deconvolution throughput (voxel-view-iterations/s) on N B200s.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0 (the LAST line of stdout; reported extras are printed before it, one short JSON line
each, and the complete record is also written to gpurun_out/bench_full_N<N>.json when that directory exists).

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): per GPU a
512x512x256 fp32 volume, 7 views, 31^3 PSFs, Efficient-Bayesian, Tikhonov lambda 0.006,
synthetic specimen data.  One step = one full iteration (7 view-steps = 14 FFT convolutions with
their fused ratio / update epilogues).  At N > 1 the global volume is N bricks of that size
(2 -> 1024x512x256, 4 -> 1024x1024x256, 8 -> 1024x1024x512 = the size of configs[2]) with the
PSF/2-wide halos pushed over NVLink peer memory every convolution: weak scaling.

value    = N_voxels(global) * views * K / device time of K iterations, inputs resident in HBM.
e2e      = the same metric through the reference-facing call -- at N = 1 literally the plugin's call
           `MVDeconvolution(views, PSFTYPE, numIterations, lambda, ...)` of spim_registration_b200/deconvolution.py
           (MVDeconvolution.java:94-211) on pinned host arrays: upload of all views, init, K iterations with their
           statistics, mask, download of psi; at N > 1 the brick runner (one such call per rank) -- host<->device copies
           inside the timed region.
roofline = dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json.
At N > 1 two more objects: `strong_scaling` (BASELINE configs[2]: ONE fixed 6-view 1024x1024x512 Optimization-II volume
split over the N GPUs, against the same volume on one GPU of the same box) and, at N = 8, `configs4` (BASELINE
configs[4]: 8 views, 2048x2048x1024, bricks of 1024x1024x512, 10 iterations).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BRICK = (256, 512, 512)      # (z, y, x) per GPU
VIEWS = 7
PSF = 31
LAMBDA = 0.006
ITER_TYPE = 2                # EFFICIENT_BAYESIAN
VARIANT_ITERS = 6            # timed iterations of a --variant-child run
METRIC = "MV deconvolution voxel-view-iters/s"      # BASELINE.json: "MV deconvolution voxel-view-iters/s at 1/2/4/8 B200"
UNIT = "voxel-view-iters/s"
KERNEL_NAMES = ["x_fwd_r2c", "y_fwd", "z_fwd_mul_inv", "y_inv", "x_inv_c2r_epilogue"]
KERNEL_ALG_FACTOR = [8, 8, 12, 8, 8]   # algorithmic bytes per launch = factor * Np
# the library times the x-inverse launches per epilogue: timer slot 4 = ratio (conv1), slot 7 = update (conv2)
XINV_SLOTS = {"x_inv_c2r_ratio": 4, "x_inv_c2r_update": 7}


def merge_xinv(kms, kcnt):
    """per-kernel (ms, launches) lists of the library -> the five sweep kernels (x-inverse = both epilogues together) plus
    the two epilogues separately"""
    kms, kcnt = list(kms), list(kcnt)
    sep = {n: (kms[i] / kcnt[i]) for n, i in XINV_SLOTS.items() if kcnt[i] > 0}
    kms[4], kcnt[4] = kms[4] + kms[7], kcnt[4] + kcnt[7]
    return kms, kcnt, sep
# the x-inverse launch also carries the fused pointwise traffic of SURVEY 8d (16 B per voxel-view-iteration):
# ratio epilogue reads img (4 N), update epilogue reads weight + psi (8 N); their writes are the sweep's own write
XINV_POINTWISE_BYTES_PER_VOXEL = 6     # average of the two launches of a view-step


def usable_cores():
    """cores this process may really use: the affinity mask, capped by the cgroup CPU quota (os.cpu_count() reports the
    whole host on shared multi-GPU boxes, and oversubscribed pocketfft pools halve the CPU baseline)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return max(1, n)


def config_dict(n_gpus):
    """the `config` object -- built the same way by both arms (ours and --impl reference) from host-side information only"""
    from spim_registration_b200 import bricks, native
    grid = bricks.grid_for(n_gpus)
    try:
        lib = native.load_library()
        fd = [int(lib.mvd_fft_size(BRICK[d] + PSF - 1, 1 if d == 2 else 0)) for d in range(3)]
    except Exception:
        fd = None
    return {"workload": workload_string(n_gpus)[0], "fft_dims_zyx": fd,
            "np_voxels_per_brick": int(np.prod([b + PSF - 1 for b in BRICK])),
            "parallelism": (f"bricks {grid[2]}x{grid[1]}x{grid[0]} (x,y,z), PSF/2 halos pushed over NVLink peer memory (NCCL batch as "
                            "fallback)") if n_gpus > 1 else "single GPU",
            "l2": "working set per convolution (>= 256 MiB real + 370 MiB spectrum) exceeds the 126 MB L2"}


def workload_string(n_gpus):
    from spim_registration_b200 import bricks
    grid = bricks.grid_for(n_gpus)
    g = tuple(BRICK[d] * grid[d] for d in range(3))
    return (f"{VIEWS}-view {g[2]}x{g[1]}x{g[0]} fp32 ({n_gpus} brick(s) of {BRICK[2]}x{BRICK[1]}x{BRICK[0]}), {PSF}^3 PSFs, "
            "Efficient-Bayesian, lambda 0.006, gen-2 semantics; step = one iteration over all views"), grid


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(shape, rank_seed=0, fast=False):
    """Synthetic specimen views for one brick; the forward blur runs on the GPU through mvd_convolve."""
    from spim_registration_b200 import native, synthetic
    lib = native.load_library()
    dev = int(os.environ.get("LOCAL_RANK", "0"))

    def blur(t, k):
        return native.convolve(t, k, 2, 0.0, device=dev, lib=lib)

    psfs = synthetic.make_psfs(VIEWS, PSF)
    if fast:
        # timing-only child runs (the direct-push variant): uniform noise shared by all views, constant weights
        rng = np.random.default_rng(11 + rank_seed)
        img = (0.05 + 0.95 * rng.random(shape, dtype=np.float32)).astype(np.float32)
        w = np.full(shape, np.float32(1.0 / VIEWS), np.float32)
        return [img] * VIEWS, [w] * VIEWS, psfs
    truth = synthetic.specimen_truth(shape, seed=2929 + rank_seed)
    imgs, ws = synthetic.make_views(truth, psfs, seed=7 + rank_seed, blur=blur)
    return imgs, ws, psfs


def cpu_reference_step(imgs, ws, psfs, psi, k1, k2, v):
    """One view-step of the oracle (SciPy pocketfft on every host core) -- the CPU restatement of the
    reference's Java path (the JVM and its jars are not available)."""
    from oracle import mvdecon_oracle as O
    p = O.DeconParams(iteration_type=ITER_TYPE, lam=LAMBDA, gen=O.GEN2)
    new, _, _ = O.view_step(psi, imgs[v], ws[v], k1[v], k2[v], p)
    return new


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path on the host cores this process may use; each timed
    step is a bounded sample (ONE view-step = 1/VIEWS of an iteration of one brick), value normalised to the metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import synthetic
    cores = usable_cores()
    O.WORKERS = cores
    shape = BRICK
    truth = synthetic.specimen_truth(shape)
    psfs = synthetic.make_psfs(VIEWS, PSF)
    nsample = min(VIEWS, 2)                          # views actually materialised for the sample
    imgs, ws = synthetic.make_views(truth, psfs[:nsample], seed=7)
    k1, k2 = O.init_kernels(psfs, ITER_TYPE)
    psi = np.full(shape, np.float32(0.05), np.float32)
    for i in range(args.warmup):
        psi = cpu_reference_step(imgs, ws, psfs, psi, k1, k2, i % nsample)
    t0 = time.perf_counter()
    for i in range(args.steps):
        psi = cpu_reference_step(imgs, ws, psfs, psi, k1, k2, i % nsample)
    dt = time.perf_counter() - t0
    nvox = int(np.prod(shape))
    value = nvox * args.steps / dt                   # one view-step = N voxel-view-iterations
    sample = (f"each timed step = ONE view-step (1/{VIEWS} of an iteration) of one {BRICK[2]}x{BRICK[1]}x{BRICK[0]} brick on the host "
              f"CPU, {args.steps} steps; CPU restatement of the reference (NumPy + SciPy pocketfft, {cores} threads); the "
              "reference JVM is not available")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def fusion_prestep_leg(shape, views, peak, lib=None, timed=None, stack_planes=None):
    """Reported extra (the callers on the input side of the hot path): raw stacks ->
    device-side affine transformation + blending weights (one ResampleK launch per view) -> weight normalisation
    (WeightNormK).  Algorithmic bytes: transform = stack read once + image and weight written once;
    normalisation = every weight read and written once + the sum image written."""
    from spim_registration_b200 import fusion
    from spim_registration_b200.deconvolution import Session
    import math
    nz, ny, nx = shape
    sz = stack_planes or max(2, int(round(nz / 2.5)))          # anisotropic raw stacks (z spacing 2.5)
    rng = np.random.default_rng(99)
    nvox = nz * ny * nx
    if timed is None:
        def timed(fn):
            t0 = time.perf_counter(); fn(); return 1e3 * (time.perf_counter() - t0)
    t_ms = []
    with Session(shape, views, ITER_TYPE, generation=2, lam=LAMBDA, lib=lib) as s:
        for v in range(views):
            stack = rng.random((sz, ny, nx), dtype=np.float32)
            a = 2.0 * math.pi * v / views
            c, si, zs = math.cos(a), math.sin(a), 2.5
            cx, cz = (nx - 1) / 2.0, (sz - 1) / 2.0 * zs
            model = fusion.AffineTransform3D([c, 0, si * zs, -(c * cx + si * cz) + cx, 0, 1, 0, 0,
                                              -si, 0, c * zs, -(-si * cx + c * cz) + cz])
            fusion.load_stack(s, stack, normalize=True)
            bl = fusion.Blending((nx, ny, sz), (-8, -8, -3), (12, 12, 12))   # EfficientBayesianBased.java:96-97, 694-696
            t_ms.append(timed(lambda: fusion.transform_view(s, v, model, (0, 0, 0), bl)))
        fusion.load_stack(s, None)
        t_norm = timed(lambda: fusion.normalize_weights(s, virtual=True, num_portions=2 * (os.cpu_count() or 1)))
    tr_ms = statistics.median(t_ms)
    tr_bytes = 4 * sz * ny * nx + 8 * nvox
    nm_bytes = (8 * views + 4) * nvox
    return {"workload": f"{views} raw stacks {nx}x{ny}x{sz} -> {nx}x{ny}x{nz} bounding box, blending border -8,-8,-3 range 12, "
                        "virtual weights", "gpu_launches": views + 1,
            "transform": {"ms_per_view": tr_ms, "voxels_per_s": nvox / (tr_ms * 1e-3), "alg_bytes": tr_bytes,
                          "gbs": tr_bytes / (tr_ms * 1e-3) / 1e9, "frac": tr_bytes / (tr_ms * 1e-3) / 1e9 / peak},
            "normalize_weights": {"ms": t_norm, "alg_bytes": nm_bytes, "gbs": nm_bytes / (t_norm * 1e-3) / 1e9,
                                  "frac": nm_bytes / (t_norm * 1e-3) / 1e9 / peak},
            "timing": "host clock around the synchronous C-ABI call (kernel launch + stream synchronise)"}


# A/B matrix of the in-tree kernel variants, measured in the same run as the headline number (rank 0, N = 1, child processes
# without torch: numpy inputs, the session C-ABI, host clock around the synchronous mvd_run and the library's own per-kernel
# CUDA events).  The first entry is the control: the default configuration in the very same harness.
VARIANTS = [
    ("default", {}),
    ("x_forward_without_tma", {"SPIM_XFWD_TMA": "0"}),             # plain-load x-forward kernel instead of the TMA-fed pipeline
    ("ieee_epilogue", {"SPIM_FAST_EPI": "0"}),                     # IEEE division / sqrt instead of the branch-free refinement
    ("column_passes_cp_async", {"SPIM_COLP": "2"}),                # y passes with one-shot cp.async staging instead of the TMA pipeline
    ("literal_constant_extension", {"SPIM_CONST_SHIFT": "0"}),     # gen-2 conv2 transforms its constant halo instead of shifting it away
]


def variant_child():
    """One configuration of the A/B matrix (environment already set by the parent); prints one JSON line."""
    from spim_registration_b200 import synthetic
    from spim_registration_b200.deconvolution import Session
    rng = np.random.default_rng(1)
    img = (0.05 + 0.95 * rng.random(BRICK, dtype=np.float32)).astype(np.float32)
    w = np.full(BRICK, np.float32(1.0 / VIEWS), np.float32)
    psfs = synthetic.make_psfs(VIEWS, PSF)
    iters = VARIANT_ITERS
    with Session(BRICK, VIEWS, ITER_TYPE, generation=2, lam=LAMBDA) as s:
        for v in range(VIEWS):
            s.set_view(v, img, w, psfs[v])
        s.init()
        s.run(2, stats=False)
        t0 = time.perf_counter()
        s.run(iters, stats=False)
        dt = time.perf_counter() - t0
        s.set_timing(True)
        s.run(2, stats=False)
        kms, kcnt, xinv_sep = merge_xinv(*s.get_timing())
        info = s.info()
        sample = s.get_psi()[::8, ::16, ::16].astype(np.float64)
        psi_ok = bool(np.isfinite(sample).all())
    per = {KERNEL_NAMES[i]: kms[i] / kcnt[i] for i in range(len(KERNEL_NAMES)) if kcnt[i] > 0}
    ms_conv = sum(per.values())
    gbs = 44 * int(info.np_voxels) / (ms_conv * 1e-3) / 1e9 if ms_conv > 0 else None
    print(json.dumps({"value": int(np.prod(BRICK)) * VIEWS * iters / dt, "ms_per_step": 1e3 * dt / iters,
                      "ms_per_conv": ms_conv, "conv_pass_gbs": gbs, "conv_pass_frac": (gbs / peaks()[0]) if gbs else None,
                      "per_kernel_ms": per, "x_inv_by_epilogue_ms": xinv_sep, "fft_dims_zyx": list(info.fft_dims), "finite": psi_ok,
                      "psi_checksum": float(sample.sum())}))


def config2_one_gpu_leg(limit_s=120.0, env_extra=None):
    """BASELINE configs[2] on ONE GPU -- the volume north_star quotes its 60 % target on: 6 views, 1024 x 1024 x 512, 31^3 PSFs,
    Optimization II (FFT size 1080 x 1080 x 560, narrow column tiles selected automatically, ~62 GB of HBM).  Same torch-free
    child as the variants; 2 + 2 iterations."""
    cmd = [sys.executable, os.path.abspath(__file__), "--variant-child", "--views", "6", "--brick", "512", "1024", "1024",
           "--iter-type", "0", "--variant-iters", "2"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit_s, env=dict(os.environ, **(env_extra or {})))
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        d = json.loads(lines[-1]) if (r.returncode == 0 and lines) else {"error": f"exit {r.returncode}: {(r.stderr or '').strip()[-200:]}"}
    except Exception as e:      # noqa: BLE001
        d = {"error": f"{type(e).__name__}: {e}"}
    d["workload"] = "6-view 1024x1024x512 fp32, 31^3 PSFs, Optimization II, one GPU, noise inputs (timing only)"
    if env_extra:
        d["env"] = env_extra
    return d


def variants_leg(budget_s=120.0, per_child_s=30.0):
    out, t_start = {}, time.perf_counter()
    for name, env in VARIANTS:
        if time.perf_counter() - t_start > budget_s - 6.0:       # a child takes about six seconds
            out[name] = {"skipped": "time budget of this extra used up"}
            continue
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--variant-child", "--views", str(VIEWS),
                   "--brick", str(BRICK[0]), str(BRICK[1]), str(BRICK[2])]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=per_child_s, env=dict(os.environ, **env))
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out[name] = json.loads(lines[-1]) if (r.returncode == 0 and lines) else {
                "error": f"exit {r.returncode}: {(r.stderr or '').strip()[-200:]}"}
        except Exception as e:      # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        out[name]["env"] = env
    ref = out.get("default", {}).get("psi_checksum")
    for name, d in out.items():      # same data, same iterations: every variant must land on the control's result
        if ref and "psi_checksum" in d:
            d["matches_default"] = bool(abs(d["psi_checksum"] - ref) <= 1e-5 * abs(ref))
    return out


def emit_extra(name, obj):
    """a reported extra as its own short line BEFORE the final line (the final line stays small enough to survive any tail)"""
    try:
        print(json.dumps({"extra": name, "data": obj}), flush=True)
    except Exception:
        pass


def strong_scaling_leg(rank, N, local, dist, iters=5):
    """BASELINE configs[2]: ONE fixed 6-view 1024x1024x512 volume, 31^3 PSFs, Optimization II, split over the N GPUs (hash
    inputs generated on the devices), against the same volume on one GPU of the same box (rank 0 alone, the others wait)."""
    import torch
    from spim_registration_b200 import bigvolume, bricks
    G = (512, 1024, 1024)
    grid = bricks.grid_for(N)
    brick = tuple(G[d] // grid[d] for d in range(3))
    out = {"workload": "6-view 1024x1024x512 fp32, 31^3 PSFs, Optimization II (BASELINE configs[2]), hash-noise inputs generated on the "
                       "devices, %d iterations timed after one warm-up" % iters, "n_gpus": N, "brick_zyx": list(brick)}
    r, meta = bigvolume.setup_runner(brick, 6, 0, rank, N, local, dist)
    ms = bigvolume.time_iterations(r, iters, dist)
    out["ms_per_iteration"] = ms
    out["value"] = int(np.prod(G)) * 6 / (ms * 1e-3)
    out["exchange"] = meta["exchange"]
    out["peak_device_bytes_per_gpu"] = bigvolume.peak_device_bytes(dist, N)
    r.close()
    del r
    torch.cuda.empty_cache()
    dist.barrier()
    one = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == 0:
        try:
            r1, meta1 = bigvolume.setup_runner(G, 6, 0, 0, 1, local, None)
            ms1 = bigvolume.time_iterations(r1, max(2, iters // 2), None)
            one[0] = ms1
            per = bigvolume.kernel_times(r1, 1)
            tot = sum(per.values())
            one[1] = 44 * meta1["np_voxels_per_brick"] / (tot * 1e-3) / 1e9 / peaks()[0]
            r1.close()
            del r1
            torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001
            out["one_gpu_error"] = f"{type(e).__name__}: {e}"
    dist.all_reduce(one)            # doubles as the barrier the other ranks wait in
    ms1 = float(one[0].item())
    if ms1 > 0:
        out["one_gpu_ms_per_iteration"] = ms1
        out["one_gpu_value"] = int(np.prod(G)) * 6 / (ms1 * 1e-3)
        out["one_gpu_conv_pass_frac"] = float(one[1].item())
        out["speedup_vs_one_gpu"] = ms1 / ms
        out["efficiency"] = ms1 / ms / N
    return out


def configs4_leg(rank, N, local, dist, iters=10):
    """BASELINE configs[4]: 8 views, 2048x2048x1024, 31^3 PSFs, Efficient-Bayesian, bricks of 1024x1024x512 on 8 GPUs, views
    handed over cell by cell with mvd_upload_region from device memory, 10 iterations (oracle spot check of the same run:
    tests/run_bricks_fullsize.py --config c5, log under profiles/)."""
    import torch
    from spim_registration_b200 import bigvolume
    brick = (512, 1024, 1024)
    r, meta = bigvolume.setup_runner(brick, 8, ITER_TYPE, rank, N, local, dist)
    ms = bigvolume.time_iterations(r, iters, dist)
    G = meta["global_zyx"]
    out = {"workload": "8-view 2048x2048x1024 fp32, 31^3 PSFs, Efficient-Bayesian, bricks of 1024x1024x512 (BASELINE configs[4]), "
                       "hash-noise inputs generated on the devices and handed over with mvd_upload_region",
           "iterations": iters, "ms_per_iteration": ms, "value": int(np.prod(G)) * 8 / (ms * 1e-3), "unit": UNIT,
           "fft_dims_zyx": meta["fft_dims_zyx"], "exchange": meta["exchange"],
           "peak_device_bytes_per_gpu": bigvolume.peak_device_bytes(dist, N), "session_device_bytes": meta["session_device_bytes"],
           "t_generate_upload_s": meta["t_generate_upload_s"], "t_init_s": meta["t_init_s"]}
    r.close()
    del r
    torch.cuda.empty_cache()
    return out


_T0 = time.perf_counter()


def tlog(msg):
    if os.environ.get("SPIM_BENCH_TRACE") and int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write(f"[bench +{time.perf_counter() - _T0:7.1f}s] {msg}\n")
        sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fusion-leg", action="store_true")
    ap.add_argument("--no-cufft-leg", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the A/B matrix of kernel variants (an extra of the default run)")
    ap.add_argument("--variant-child", action="store_true", help="internal: one configuration of the A/B matrix")
    ap.add_argument("--no-big-legs", action="store_true", help="N > 1: skip the strong-scaling (configs[2]) and configs[4] legs")
    ap.add_argument("--fast-inputs", action="store_true", help="internal: noise inputs instead of the synthetic specimen")
    ap.add_argument("--skip-e2e", action="store_true", help="internal: timing-only child runs")
    ap.add_argument("--fusion-leg-only", action="store_true", help="internal: run the fusion pre-step leg and print its JSON")
    ap.add_argument("--brick", type=int, nargs=3, default=None, help="per-GPU brick (z y x), default 256 512 512")
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--iter-type", type=int, default=None, help="PSFTYPE ordinal (0 Optimization II ... 3 independent), default 2")
    ap.add_argument("--variant-iters", type=int, default=None)
    args = ap.parse_args()
    global BRICK, VIEWS, ITER_TYPE, VARIANT_ITERS
    if args.iter_type is not None:
        ITER_TYPE = args.iter_type
    if args.variant_iters:
        VARIANT_ITERS = args.variant_iters
    if args.brick:
        BRICK = tuple(args.brick)
    if args.views:
        VIEWS = args.views
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    if args.fusion_leg_only:
        print(json.dumps(fusion_prestep_leg(BRICK, VIEWS, peaks()[0])))
        return
    if args.variant_child:
        variant_child()
        return

    tlog("start")
    import torch
    from spim_registration_b200 import build as b
    tlog("torch imported")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        b.build_cuda_library()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        tlog("process group up")
    if world != args.gpus:
        if rank == 0:
            sys.stderr.write(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE\n")
    N = world

    from spim_registration_b200 import native
    from spim_registration_b200.deconvolution import Session
    from spim_registration_b200 import bricks

    grid = bricks.grid_for(N)                        # bricks along (z, y, x)
    coords = bricks.rank_coords(rank, grid)
    gshape = tuple(BRICK[d] * grid[d] for d in range(3))
    imgs, ws, psfs = make_inputs(BRICK, rank_seed=rank, fast=args.fast_inputs)
    tlog("inputs generated")
    nvox_brick = int(np.prod(BRICK))
    nvox_global = nvox_brick * N

    # pinned host copies (the hand-over format of the reference: materialised fp32 arrays)
    pin_img = [torch.from_numpy(a).pin_memory() for a in imgs]
    pin_w = [torch.from_numpy(a).pin_memory() for a in ws]
    pin_out = torch.empty(BRICK, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(t.numel() * 4 for t in pin_img + pin_w) + sum(p.size * 4 for p in psfs)
    d2h_bytes = pin_out.numel() * 4
    tlog("pinned")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def new_runner():
        r = bricks.BrickRunner(BRICK, VIEWS, ITER_TYPE, generation=2, lam=LAMBDA, device=local, rank=rank,
                               world=N, grid=grid, dist=dist)
        return r

    def upload(r):
        for v in range(VIEWS):
            r.session.set_view_ptr(v, pin_img[v].data_ptr(), pin_w[v].data_ptr(), psfs[v])

    # ---------------- device-resident throughput -------------------------------------------------
    runner = new_runner()
    upload(runner)
    tlog("uploaded")
    runner.init()
    tlog("init done")
    info = runner.session.info()
    exchange_path = ("direct push over peer memory (mvd_p2p_*)" if runner.use_p2p else
                     "single-launch pack / unpack around one NCCL batch" if runner.use_pack else
                     "slab copies around one NCCL batch") if N > 1 else None
    np_brick = int(info.np_voxels)
    stream = torch.cuda.ExternalStream(runner.session.stream(), device=torch.device("cuda", local))
    for _ in range(args.warmup):
        runner.run(1)
    barrier()
    tlog("warmup done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    runner.run(args.steps)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    tlog("timed region done")
    clocks = sampler.stop() if rank == 0 else None
    value = nvox_global * VIEWS * args.steps / (ms * 1e-3)
    launches = 10 * VIEWS * args.steps * 1 + runner.extra_launches_per_iteration() * args.steps

    # ---------------- per-kernel timing for the roofline (separate run, events around every launch) -----
    runner.session.set_timing(True)
    runner.use_graph = False          # brick mode: eager launches here, so that every kernel is bracketed by its events
    runner.run(max(2, min(args.steps, 5)))
    kms, kcnt, xinv_sep = merge_xinv(*runner.session.get_timing())
    runner.session.set_timing(False)
    peak, peak_src = peaks()
    per_kernel = {}
    for i, name in enumerate(KERNEL_NAMES):
        if kcnt[i] > 0:
            avg_ms = kms[i] / kcnt[i]
            alg = KERNEL_ALG_FACTOR[i] * np_brick + (XINV_POINTWISE_BYTES_PER_VOXEL * nvox_brick if i == 4 else 0)
            per_kernel[name] = {"avg_ms": avg_ms, "launches": int(kcnt[i]), "alg_bytes": alg,
                                "gbs": alg / (avg_ms * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["avg_ms"] * per_kernel[k]["launches"]) if per_kernel else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = None
    conv_pass = None
    if dom:
        a = per_kernel[dom]["gbs"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": per_kernel[dom]["alg_bytes"], "avg_ms": per_kernel[dom]["avg_ms"]}
        tot_ms = sum(per_kernel[k]["avg_ms"] for k in per_kernel)
        a2 = 44 * np_brick / (tot_ms * 1e-3) / 1e9
        conv_pass = {"achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak, "ms_per_conv": tot_ms, "x_inv_by_epilogue_ms": xinv_sep,
                     "alg_bytes": 44 * np_brick, "per_kernel": per_kernel,
                     "note": "44*Np per FFT-convolution pass (SURVEY 8d), Np = prod(n + k - 1); sum of the five kernels' average durations"}
        # whole view-step: two passes + 16 B/voxel pointwise = B_vvi * N
        vs_bytes = 88 * np_brick + 16 * nvox_brick
        a3 = vs_bytes / (2 * tot_ms * 1e-3) / 1e9
        view_step = {"achieved": a3, "peak": peak, "unit": "GB/s", "frac": a3 / peak, "alg_bytes": vs_bytes}
    runner.close()
    del runner
    tlog("kernel timing done")

    # ---------------- end to end through the reference-facing call -------------------------------------
    barrier()
    t_e2e, e2e_value, e2e_note = None, None, None
    if not args.skip_e2e:
        if N == 1:
            # the plugin's own call: MVDeconInput.add(new MVDeconFFT(image, weight, kernel, ...)) per view, then the
            # MVDeconvolution constructor, which runs all iterations (MVDeconvolution.java:94-211), then getPsi()
            from spim_registration_b200.deconvolution import MVDeconFFT, MVDeconInput, MVDeconvolution, PSFTYPE
            out_np = pin_out.numpy()
            t0 = time.perf_counter()
            views = MVDeconInput()
            for v in range(VIEWS):
                views.add(MVDeconFFT(pin_img[v].numpy(), pin_w[v].numpy(), psfs[v], None, (local,), False, None, False))
            decon = MVDeconvolution(views, PSFTYPE(ITER_TYPE), args.steps, LAMBDA, 1.0, 0, "bench")
            decon.getPsi(out_np)
            torch.cuda.synchronize()
            t_e2e = time.perf_counter() - t0
            del decon, views
            e2e_note = ("MVDeconInput.add(MVDeconFFT(...)) x views + MVDeconvolution(views, PSFTYPE, K, lambda, ...) + getPsi(): upload of "
                        "all views from pinned memory, init, K iterations with per-view statistics, mask, download of psi")
        else:
            t0 = time.perf_counter()
            r2 = new_runner()
            upload(r2)
            r2.init()
            r2.run(args.steps)
            r2.finish()
            r2.session.get_psi_ptr(pin_out.data_ptr())
            torch.cuda.synchronize()
            t_e2e = max_over_ranks(time.perf_counter() - t0)
            r2.close()
            e2e_note = ("per rank: brick session create + upload of all views from pinned memory + init (average all-reduced, peers "
                        "connected) + K iterations + mask + download of the brick of psi")
        e2e_value = nvox_global * VIEWS * args.steps / t_e2e
        tlog("e2e done")

    # ---------------- N > 1: strong scaling on the configs[2] volume, configs[4] at N = 8 ----------------------------------
    strong, configs4 = None, None
    if N > 1 and not args.no_big_legs and tuple(BRICK) == (256, 512, 512):
        try:
            strong = strong_scaling_leg(rank, N, local, dist)
        except Exception as e:      # noqa: BLE001
            strong = {"error": f"{type(e).__name__}: {e}"}
        barrier()
        tlog("strong-scaling leg done")
        if N == 8:
            try:
                configs4 = configs4_leg(rank, N, local, dist)
            except Exception as e:      # noqa: BLE001
                configs4 = {"error": f"{type(e).__name__}: {e}"}
            barrier()
            tlog("configs[4] leg done")

    # ---------------- CPU baseline: the oracle on a bounded sample (rank 0, N = 1 only) --------------------
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        from oracle import mvdecon_oracle as O
        O.WORKERS = usable_cores()
        k1, k2 = O.init_kernels(psfs, ITER_TYPE)
        psi = np.full(BRICK, np.float32(info.avg), np.float32)
        nsteps = 2
        t0 = time.perf_counter()
        for v in range(nsteps):
            psi = cpu_reference_step(imgs, ws, psfs, psi, k1, k2, v)
        dt = time.perf_counter() - t0
        cpu = {"value": nvox_brick * nsteps / dt, "unit": UNIT, "cores": usable_cores(), "kind": "port",
               "sample": f"{nsteps} view-steps (of {VIEWS * args.steps}) of the same {BRICK[2]}x{BRICK[1]}x{BRICK[0]} "
                         "workload; CPU restatement of the reference (NumPy + SciPy pocketfft, all usable cores); "
                         "reference JVM unavailable"}

    # ---------------- reported extras (rank 0, N = 1; child processes, so that none of them can take the line with it) ------
    # All extras together respect one wall-clock deadline counted from the start of this process (SPIM_BENCH_EXTRAS_S, default
    # 230 s): whatever does not fit is reported as skipped, so that the default run ends within about four minutes.
    deadline = _T0 + float(os.environ.get("SPIM_BENCH_EXTRAS_S", "230"))

    def time_left():
        return deadline - time.perf_counter()

    skipped = {"skipped": "time budget of the extras used up"}

    # the same convolution built on cuFFT (comparison point only, separate executable)
    cufft_leg = None
    if rank == 0 and N == 1 and not args.no_cufft_leg:
        try:
            exe = b.build_cufft_comparison() if time_left() > 15 else None
            if exe:
                fd = [int(v) for v in info.fft_dims]
                cmd = [exe, str(BRICK[0]), str(BRICK[1]), str(BRICK[2]), str(PSF), str(fd[0]), str(fd[1]), str(fd[2]), "20"]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=max(15.0, min(120.0, time_left())))
                out = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                cufft_leg = json.loads(out[-1]) if out else {"error": f"exit {r.returncode}: {(r.stderr or '').strip()[-200:]}"}
                if conv_pass and "ms_per_conv" in cufft_leg:
                    cufft_leg["ours_ms_per_conv"] = conv_pass["ms_per_conv"]
                    cufft_leg["speedup_ours_vs_cufft"] = cufft_leg["ms_per_conv"] / conv_pass["ms_per_conv"]
            else:
                cufft_leg = dict(skipped) if time_left() <= 15 else {"error": "comparison binary not built"}
        except Exception as e:      # noqa: BLE001
            cufft_leg = {"error": f"{type(e).__name__}: {e}"}
        tlog("cufft leg done")

    # A/B matrix of the in-tree kernel variants
    variants = None
    if rank == 0 and N == 1 and not args.no_variants:
        try:
            variants = variants_leg(budget_s=min(120.0, time_left() - 35.0))      # keep room for the configs[2] leg
        except Exception as e:      # noqa: BLE001
            variants = {"error": f"{type(e).__name__}: {e}"}
        tlog("variants done")

    # the north_star target volume (configs[2]) on this one GPU
    config2_leg = None
    if rank == 0 and N == 1 and not args.no_variants and tuple(BRICK) == (256, 512, 512):
        config2_leg = config2_one_gpu_leg(limit_s=max(30.0, min(120.0, time_left() + 30.0))) if time_left() > 10 else dict(skipped)
        if time_left() > 40 and isinstance(config2_leg, dict) and "value" in config2_leg:
            # the same with 16-column tiles forced (one 138 KB tile per SM on the 1080-long axes): what the narrow tiles buy
            config2_leg["with_16_column_tiles"] = config2_one_gpu_leg(limit_s=max(30.0, min(90.0, time_left())),
                                                                      env_extra={"SPIM_COL_NARROW": "0"})
        tlog("configs[2] leg done")

    # the device-side fusion pre-step
    fusion_leg = None
    if rank == 0 and N == 1 and not args.no_fusion_leg:
        if time_left() > 20:
            try:
                cmd = [sys.executable, os.path.abspath(__file__), "--fusion-leg-only", "--views", str(VIEWS),
                       "--brick", str(BRICK[0]), str(BRICK[1]), str(BRICK[2])]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=max(30.0, min(120.0, time_left() + 30.0)))
                out = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                fusion_leg = json.loads(out[-1]) if (r.returncode == 0 and out) else {
                    "error": f"child exited with {r.returncode}: {(r.stderr or '').strip()[-300:]}"}
            except Exception as e:      # noqa: BLE001
                fusion_leg = {"error": f"{type(e).__name__}: {e}"}
        else:
            fusion_leg = dict(skipped)
        tlog("fusion leg done")

    if rank == 0:
        extras = {"fusion_prestep": fusion_leg, "cufft_comparison": cufft_leg, "variants": variants,
                  "configs2_one_gpu": config2_leg, "roofline_conv_pass_per_kernel": (conv_pass or {}).get("per_kernel"),
                  "strong_scaling": strong, "configs4": configs4}
        for k, v in extras.items():
            if v is not None:
                emit_extra(k, v)
        cfg = config_dict(N)
        cfg["fft_dims_zyx"] = [int(x) for x in info.fft_dims]
        cfg["np_voxels_per_brick"] = np_brick
        brief = lambda d, keys: None if not isinstance(d, dict) else {k: d.get(k) for k in keys if k in d}      # noqa: E731
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": None if e2e_value is None else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes / args.steps,
                "d2h_bytes_per_step": d2h_bytes / args.steps, "seconds": t_e2e, "note": e2e_note},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_conv_pass": None if not conv_pass else {k: v for k, v in conv_pass.items() if k != "per_kernel"},
            "roofline_view_step": view_step if dom else None,
            "cpu_baseline": cpu,
            "exchange_path": exchange_path,
            "strong_scaling": brief(strong, ("value", "ms_per_iteration", "one_gpu_value", "one_gpu_ms_per_iteration",
                                             "one_gpu_conv_pass_frac", "speedup_vs_one_gpu", "efficiency", "exchange", "error")),
            "configs4": brief(configs4, ("value", "ms_per_iteration", "iterations", "peak_device_bytes_per_gpu", "exchange", "error")),
            "configs2_one_gpu": brief(config2_leg, ("value", "ms_per_step", "ms_per_conv", "conv_pass_frac", "per_kernel_ms", "error", "skipped")),
            "extras_printed_before_this_line": [k for k, v in extras.items() if v is not None],
        }
        text = json.dumps(line)
        try:
            d_ = os.path.join(ROOT, "gpurun_out")
            if os.path.isdir(d_):
                with open(os.path.join(d_, f"bench_full_N{N}.json"), "w") as f:
                    f.write(json.dumps(dict(line, extras=extras)) + "\n")
        except Exception:
            pass
        print(text)
    sys.stdout.flush()
    sys.stderr.flush()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
        tlog("exiting")
        # hard exit: tearing down NCCL communicators that were captured into CUDA graphs can block for minutes
        os._exit(0)


if __name__ == "__main__":
    main()
