#!/bin/bash
# fused halo push A/B on N GPUs:  gpurun --gpus N -- bash profiles/r2_fuse_ab.sh N
set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 python -m pytest tests/test_multigpu_nccl.py -m gpu -q -k "over_nccl or direct_push or one_thread" 2>&1 | tail -3
timeout 420 $TR --master-port 29541 tests/run_bricks_fullsize.py --config c3 --json gpurun_out/r2_fuse_c3_N$N.json > gpurun_out/r2_fuse_c3_N$N.txt 2>&1
grep -E "fullsize|FULLSIZE|Error" gpurun_out/r2_fuse_c3_N$N.txt | tail -6
for f in 1 0 1 0; do
    SPIM_BRICK_FUSE=$f timeout 300 $TR --master-port 2957$f bench.py --gpus $N --steps 20 --warmup 3 --no-big-legs > gpurun_out/r2_fuse${f}_bench_N$N.txt 2>/dev/null
    python - $f $N <<'PY'
import json, sys
f, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/r2_fuse{f}_bench_N{n}.txt").read().strip().splitlines()[-1])
    print(f"fuse={f} N={n}: ms_per_step {d['ms_per_step']:.3f} value {d['value']/1e9:.1f} G vvi/s  conv pass {d['roofline_conv_pass']['frac']:.4f}")
except Exception as e:
    print("fuse", f, "bench:", e)
PY
done
