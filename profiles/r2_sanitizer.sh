#!/bin/bash
# compute-sanitizer on hardware: memcheck on the parity subset, racecheck + synccheck on its quick form.
#   gpurun -- bash profiles/r2_sanitizer.sh      -> gpurun_out/r2_sanitizer_{memcheck,racecheck,synccheck}.txt
set -u
for tool in memcheck racecheck synccheck; do
    arg=""; [ $tool != memcheck ] && arg="quick"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tests/sanitizer_subset.py $arg > gpurun_out/r2_sanitizer_$tool.txt 2>&1
    echo "exit $?" >> gpurun_out/r2_sanitizer_$tool.txt
    echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_SUBSET_OK|exit " gpurun_out/r2_sanitizer_$tool.txt | tail -4
done
