#!/bin/bash
# ncu --set full (with source) of one view-step (10 launches) of the hot path at the bench brick size.
#   bash profiles/r2_ncu.sh <tag> [ENV=VALUE ...]      -> gpurun_out/<tag>.ncu-rep
set -u
tag=$1; shift
env "$@" timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o gpurun_out/$tag -f python profiles/prof_run.py 1 256 512 512 31 > gpurun_out/$tag.log 2>&1
tail -3 gpurun_out/$tag.log
