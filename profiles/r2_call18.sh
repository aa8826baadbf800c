TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29545 tests/run_bricks_fullsize.py --config custom --brick 512 1024 1024 --views 8 --iter-type 2 --check-steps 2 --iters 3 --json gpurun_out/r2_c5brick_2gpu.json > gpurun_out/r2_c5brick_2gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r2_c5brick_2gpu.txt
grep -E "fullsize|FULLSIZE|exit |Error" gpurun_out/r2_c5brick_2gpu.txt | tail -8
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_c5brick_2gpu.json"))
    print({k: d.get(k) for k in ("value", "ms_per_iteration", "peak_device_bytes_per_gpu", "exchange", "fft_dims_zyx", "ok")})
except Exception as e:
    print(e)
PY
