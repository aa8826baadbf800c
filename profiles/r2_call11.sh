t0=$SECONDS
SPIM_BENCH_TRACE=1 python bench.py > gpurun_out/r2_bench_default.txt 2> gpurun_out/r2_bench_default.err; echo "bench exit $? in $((SECONDS - t0)) s"
tail -1 gpurun_out/r2_bench_default.txt | cut -c1-2500
grep -E "bench \+" gpurun_out/r2_bench_default.err | tail -14
t0=$SECONDS
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.txt 2>&1; echo "reference exit $? in $((SECONDS - t0)) s"; tail -1 gpurun_out/r2_bench_reference.txt | cut -c1-900
t0=$SECONDS
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tests/sanitizer_subset.py quick > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "exit $?" >> gpurun_out/r2_sanitizer_racecheck.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_SUBSET_OK|exit " gpurun_out/r2_sanitizer_racecheck.txt | tail -3; echo "racecheck $((SECONDS - t0)) s"
