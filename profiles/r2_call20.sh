set -u
N=8; tag=r2c_mg8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
t0=$SECONDS
timeout 420 python -m pytest tests/test_multigpu_nccl.py -m gpu -q -k "over_nccl or direct_push or one_thread" > gpurun_out/${tag}_tests.txt 2>&1; tail -2 gpurun_out/${tag}_tests.txt
echo "[call20] tests $((SECONDS - t0)) s"; t0=$SECONDS
timeout 600 $TR --master-port 29542 tests/run_bricks_fullsize.py --config c5 --json gpurun_out/${tag}_c5.json > gpurun_out/${tag}_c5.txt 2>&1
grep -E "fullsize|FULLSIZE|Error" gpurun_out/${tag}_c5.txt | tail -6
echo "[call20] c5 $((SECONDS - t0)) s"; t0=$SECONDS
SPIM_BENCH_TRACE=1 timeout 600 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench.txt 2> gpurun_out/${tag}_bench.err
python - $tag <<'PY'
import json, sys
t = sys.argv[1]
d = json.load(open(f"gpurun_out/{t}_c5.json")); print("c5:", {k: d.get(k) for k in ("value", "ms_per_iteration", "peak_device_bytes_per_gpu", "ok")})
d = json.loads(open(f"gpurun_out/{t}_bench.txt").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "strong", d["strong_scaling"], "configs4", d["configs4"])
PY
echo "[call20] bench $((SECONDS - t0)) s"
