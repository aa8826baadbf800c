#!/bin/bash
# A/B matrix of the kernel variants that are in tree behind environment switches (run on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/experiments.sh > gpurun_out/experiments.txt 2>&1'
# Quickest A/B of a single switch (6 s, no torch):  env SPIM_PDL=1 python bench.py --variant-child   -> one JSON line with the
# throughput, the per-kernel CUDA-event times and a result checksum; bench.py's own `variants` extra runs the whole matrix this way.
# Every line: the switch setting, whole-iteration throughput and the per-kernel CUDA-event averages of bench.py.
# A variant is only worth adopting if the parity subset below passes with it.
set -u
summ='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d["roofline_conv_pass"]
print("  value %.4g  conv %.3f ms  frac %.3f " % (d["value"], c["ms_per_conv"], c["frac"]), {k: round(v["avg_ms"], 3) for k, v in c["per_kernel"].items()})'
run() {   # run "<env assignments>"
    echo "== $1"
    env $1 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --no-cufft-leg --no-fusion-leg 2>/dev/null | python -c "$summ" || echo "  FAILED"
}
check() { # parity subset under a variant
    echo "== parity under: $1"
    env $1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_all_extensions or golden or radix_path or config1 or randomized" 2>&1 | tail -1
}
run "SPIM_COLP=2"                                   # default: one-shot cp.async tile staging
run "SPIM_COLP=2 SPIM_REGCAP=1 SPIM_THREADS_COL=256" # 80 registers, 24 resident warps
run "SPIM_COLP=2 SPIM_THREADS_COL=160"
run "SPIM_COLP=3"                                   # TMA tensor-map pipeline, 2 consumer groups
run "SPIM_COLP=3 SPIM_THREADS_COLT=352"
run "SPIM_COLP=3 SPIM_TMAP=0"                       # per-row bulk copies (known slow)
run "SPIM_COLP=2 SPIM_COLP_Y=3"                     # hybrid: TMA pipeline for the 72 KB y tiles, default for the z pass
run "SPIM_COLP=2 SPIM_COLP_Y=3 SPIM_THREADS_COLT=352"
run "SPIM_COLP=4"                                   # warp-private columns
run "SPIM_COLP=2 SPIM_COLP_Z=4"                     # warp-private columns for the small z tiles only
run "SPIM_COLP=2 SPIM_KSTAGE=1"
run "SPIM_FAST_EPI=0"                               # IEEE intrinsics instead of the (default) branch-free MUFU-seeded division / sqrt
run "SPIM_COLP_Y=3"
run "SPIM_PDL=1"                                    # programmatic dependent launch: tail of one sweep overlaps the ramp-up of the next
run "SPIM_SERPENTINE=1"                             # y-forward / x-inverse sweeps reversed: start on the predecessor's L2 leftovers
run "SPIM_PDL=1 SPIM_SERPENTINE=1"
run "SPIM_XPLAN_ASC=1"                               # x plan smallest radix first
run "SPIM_XPLAN_ASC=1 SPIM_XINV_R0=1"                # + register-lean update kernel (80 registers, 6 blocks per SM)
run "SPIM_COL_LEAN=1"                               # z pass from the radix <= 8 instantiation (80 registers, 6 blocks)
run "SPIM_REGCAP=2"                                 # y tiles: 3 x 192 threads
run "SPIM_REGCAP=3"                                 # z tiles: 6 x 128 threads
run "SPIM_PDL=1 SPIM_COL_NARROW=1"
run "SPIM_COL_NARROW=1"                             # 8-column tiles (six 36 KB y tiles / eleven 18 KB z tiles per SM) on the C2 sizes
run "SPIM_THREADS_XFWD=128"
run "SPIM_THREADS_XFWD=256"
run "SPIM_THREADS_XINV=192"
check "SPIM_FAST_EPI=0"
check "SPIM_PDL=1"
check "SPIM_COLP=3"
check "SPIM_COLP=2 SPIM_COLP_Y=3"
check "SPIM_COLP=4"
check "SPIM_COLP=2 SPIM_REGCAP=1 SPIM_THREADS_COL=256"
# fusion pre-step (round-2 first run on hardware): parity, then one ncu pass over its kernels
echo "== fusion pre-step GPU tests"
timeout 600 python -m pytest tests/test_zz_gpu_fusion.py -x -q 2>&1 | tail -3

# written after the last hardware run of round 1: narrow column tiles, then a race check of the hot-path kernels
echo "== narrow-tile GPU tests"
timeout 600 python -m pytest tests/test_zzz_gpu_narrow_tiles.py -x -q 2>&1 | tail -3
echo "== racecheck (small parity subset)"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "conv_all_extensions or golden" 2>&1 | tail -15
# multi-GPU (run with gpurun --gpus 2): direct halo push over peer memory, then its timing against the NCCL paths
#   python -m pytest tests/test_multigpu_nccl.py -q
#   for e in "SPIM_BRICK_P2P=0" "SPIM_BRICK_P2P=1" "SPIM_BRICK_PACK=0"; do env $e python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 2 --steps 5 --warmup 3; done
# cuFFT comparison point (separate executable, comparison only): timing, then its kernels under ncu
#   profiles/microbench/cufft_conv 256 512 512 31 288 560 560 20
#   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 profiles/microbench/cufft_conv 256 512 512 31 288 560 560 2
