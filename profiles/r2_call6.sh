timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_pruned.txt 2>&1; tail -3 gpurun_out/r2_gputests_pruned.txt
AB_C3_CFGS="SPIM_NOP=2|SPIM_PDL=0" bash profiles/r2_ab.sh r2_ab_pruned "SPIM_NOP=2" "SPIM_PDL=0" "SPIM_XFWD_TMA=0" "SPIM_COLP=2" "SPIM_FAST_EPI=0"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/sanitizer_subset.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "exit $?" >> gpurun_out/r2_sanitizer_memcheck.txt
grep -E "ERROR SUMMARY|SANITIZER_SUBSET_OK|exit " gpurun_out/r2_sanitizer_memcheck.txt | tail -3
