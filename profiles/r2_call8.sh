#!/bin/bash
# The round's multi-GPU evidence in one call:   gpurun --gpus N -- bash profiles/r2_call8.sh N [quick]
#   1. tests/test_multigpu_nccl.py on N GPUs (NCCL pack path, direct push over peer memory, one process / thread per device)
#   2. BASELINE configs[2] geometry on N bricks: one full iteration, straddling sub-bricks recomputed by the oracle
#   3. (N = 8) BASELINE configs[4]: 8 views 2048x2048x1024, oracle check of the first view-steps, 10 iterations timed
#   4. halo-exchange timing alone (direct push, x pieces widened to 16-byte groups or not)
#   5. bench.py --gpus N (weak scaling; with the strong-scaling / configs[4] legs unless "quick")
# Logs -> gpurun_out/<tag>_*.txt|json (copied to profiles/r2/ afterwards).
set -u
N=${1:-8}; mode=${2:-full}; tag=${3:-r2_mg$N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${tag}_tests.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${tag}_tests.txt 2>&1
t0=$SECONDS
timeout 420 python -m pytest tests/test_multigpu_nccl.py -m gpu -v -rA -s >> gpurun_out/${tag}_tests.txt 2>&1
echo "exit $?" >> gpurun_out/${tag}_tests.txt
grep -E "PASSED|FAILED|ERROR|passed|failed|exit " gpurun_out/${tag}_tests.txt | tail -8
echo "[call8] tests $((SECONDS - t0)) s"; t0=$SECONDS
timeout 420 $TR --master-port 29541 tests/run_bricks_fullsize.py --config c3 --json gpurun_out/${tag}_c3.json > gpurun_out/${tag}_c3.txt 2>&1
echo "exit $?" >> gpurun_out/${tag}_c3.txt
grep -E "fullsize|FULLSIZE|exit |Error|error" gpurun_out/${tag}_c3.txt | tail -8
echo "[call8] c3 check $((SECONDS - t0)) s"; t0=$SECONDS
if [ "$N" = 8 ]; then
    timeout 600 $TR --master-port 29542 tests/run_bricks_fullsize.py --config c5 --json gpurun_out/${tag}_c5.json > gpurun_out/${tag}_c5.txt 2>&1
    echo "exit $?" >> gpurun_out/${tag}_c5.txt
    grep -E "fullsize|FULLSIZE|exit |Error|error" gpurun_out/${tag}_c5.txt | tail -8
    python - $tag <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/{sys.argv[1]}_c5.json"))
    print("c5:", {k: d.get(k) for k in ("value", "ms_per_iteration", "peak_device_bytes_per_gpu", "exchange", "fft_dims_zyx", "t_generate_upload_s", "t_init_s", "ok")})
except Exception as e:
    print("c5 json:", e)
PY
    echo "[call8] c5 $((SECONDS - t0)) s"; t0=$SECONDS
fi
for xp in 1 0; do
    SPIM_BRICK_XPAD=$xp timeout 200 $TR --master-port 2955$xp tests/run_exchange_timing.py 2>&1 | grep EXCHANGE_TIMING | tee -a gpurun_out/${tag}_exchange.txt
done
echo "[call8] exchange timing $((SECONDS - t0)) s"; t0=$SECONDS
extra=""; [ "$mode" = quick ] && extra="--no-big-legs"
SPIM_BENCH_TRACE=1 timeout 600 $TR --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 $extra > gpurun_out/${tag}_bench.txt 2> gpurun_out/${tag}_bench.err
echo "exit $?" >> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.txt; grep -E "bench \+|exit |Error" gpurun_out/${tag}_bench.err | tail -12
echo "[call8] bench $((SECONDS - t0)) s"
