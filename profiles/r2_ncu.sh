#!/bin/bash
# ncu --set full (with source) of one view-step (10 launches) of the hot path at the bench brick size; the report stays on
# the box (it exceeds what gpurun copies back), its pages come back as compressed CSV:
#   bash profiles/r2_ncu.sh <tag> [ENV=VALUE ...]   -> gpurun_out/<tag>_raw.csv.gz (all launches, every metric)
#                                                        gpurun_out/<tag>_src_<i>.csv.gz (SASS-level page of launch i = 0..9)
set -u
tag=$1; shift
rep=/tmp/$tag.ncu-rep
env "$@" timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o /tmp/$tag -f python profiles/prof_run.py 1 256 512 512 31 > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log
ncu -i $rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/${tag}_raw.csv.gz
for i in ${NCU_SRC_LAUNCHES:-0 4 9}; do
    ncu -i $rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 2>/dev/null | gzip -9 > gpurun_out/${tag}_src_$i.csv.gz
done
ls -la gpurun_out/${tag}_* | awk '{print $5, $9}'
