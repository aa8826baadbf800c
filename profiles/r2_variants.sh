#!/bin/bash
# Round 2, first hardware call: the complete A/B matrix of the in-tree kernel variants (every row of bench.py's VARIANTS)
# on configs[1] (7-view 512x512x256) and the decisive combinations on the configs[2] volume (6-view 1024x1024x512, one GPU),
# one JSON line per run into gpurun_out/r2_variants.jsonl -- nothing is lost in a stdout tail this time.
set -u
out=gpurun_out/r2_variants.jsonl
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_variants_gpu.txt 2>&1
python - <<'PY' >> gpurun_out/r2_variants_gpu.txt 2>&1
import bench, json, os, subprocess, sys, time
out = open("gpurun_out/r2_variants.jsonl", "a")
def child(name, env, extra):
    cmd = [sys.executable, "bench.py", "--variant-child"] + extra
    t0 = time.time()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=150, env=dict(os.environ, **env))
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        d = json.loads(lines[-1]) if (r.returncode == 0 and lines) else {"error": f"exit {r.returncode}: {(r.stderr or '').strip()[-300:]}"}
    except Exception as e:
        d = {"error": repr(e)}
    d.update(name=name, env=env, args=extra, wall_s=round(time.time() - t0, 1))
    out.write(json.dumps(d) + "\n"); out.flush()
    print(name, d.get("conv_pass_frac"), d.get("per_kernel_ms"), d.get("error"), flush=True)
c1 = ["--views", "7", "--brick", "256", "512", "512"]
for name, env in bench.VARIANTS:
    child("c1/" + name, env, c1)
c3 = ["--views", "6", "--brick", "512", "1024", "1024", "--iter-type", "0", "--variant-iters", "2"]
for name, env in [("default", {}), ("pdl", {"SPIM_PDL": "1"}), ("pdl_tma_y", {"SPIM_PDL": "1", "SPIM_COLP_Y": "3"}),
                  ("x160_lean", {"SPIM_THREADS_XFWD": "160", "SPIM_THREADS_XINV": "160", "SPIM_XPLAN_ASC": "1", "SPIM_XINV_R0": "1"}),
                  ("serpentine", {"SPIM_SERPENTINE": "1"}), ("xfwd_256", {"SPIM_THREADS_XFWD": "256"})]:
    child("c3/" + name, env, c3)
PY
