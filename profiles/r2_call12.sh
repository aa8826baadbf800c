t0=$SECONDS
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tests/sanitizer_subset.py quick > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "exit $?" >> gpurun_out/r2_sanitizer_racecheck.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_SUBSET_OK|exit " gpurun_out/r2_sanitizer_racecheck.txt | tail -3; echo "racecheck $((SECONDS - t0)) s"; t0=$SECONDS
# full-size single-GPU parity: configs[1] (7 views, 512x512x256, Efficient-Bayesian) one full iteration; configs[3] geometry
# (6 views, 1024x1024x512, independent + Tikhonov + per-view weights) one full iteration
timeout 500 python tests/run_bricks_fullsize.py --config custom --brick 256 512 512 --views 7 --iter-type 2 --json gpurun_out/r2_fullsize_c2_1gpu.json > gpurun_out/r2_fullsize_c2_1gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r2_fullsize_c2_1gpu.txt
grep -E "fullsize|FULLSIZE|exit |Error" gpurun_out/r2_fullsize_c2_1gpu.txt | tail -5; echo "c2 $((SECONDS - t0)) s"; t0=$SECONDS
timeout 700 python tests/run_bricks_fullsize.py --config custom --brick 512 1024 1024 --views 6 --iter-type 3 --check-steps 6 --json gpurun_out/r2_fullsize_c4_1gpu.json > gpurun_out/r2_fullsize_c4_1gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r2_fullsize_c4_1gpu.txt
grep -E "fullsize|FULLSIZE|exit |Error" gpurun_out/r2_fullsize_c4_1gpu.txt | tail -5; echo "c4 $((SECONDS - t0)) s"
