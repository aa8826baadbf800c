#!/bin/bash
# Round-2 evidence for the binary that ships:   gpurun -- bash profiles/r2_final_profile.sh
#   1. ncu launch list (gpu__time_duration) of a short bench.py run      -> gpurun_out/r2f_launches.csv
#   2. ncu --set full of one view-step (10 launches) with SASS pages     -> gpurun_out/r2f_default_raw.csv.gz, _src_<i>.csv.gz
set -u
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline --no-fusion-leg --no-cufft-leg > gpurun_out/r2f_launches_bench.log 2>&1
tail -c 600 gpurun_out/r2f_launches_bench.log | head -c 300; echo
grep -c "kernel_entry" gpurun_out/r2f_launches.csv
NCU_SRC_LAUNCHES="0 2 4 9" bash profiles/r2_ncu.sh r2f_default
