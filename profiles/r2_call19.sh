t0=$SECONDS
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_final.txt 2>&1; tail -3 gpurun_out/r2_gputests_final.txt; echo "gpu suite $((SECONDS - t0)) s"
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
t0=$SECONDS
python bench.py > gpurun_out/r2_bench_default.txt 2> gpurun_out/r2_bench_default.err; echo "bench exit $? in $((SECONDS - t0)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default.txt").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "pass", d["roofline_conv_pass"]["frac"], "c2_one_gpu", (d.get("configs2_one_gpu") or {}).get("conv_pass_frac"), "cpu", d["cpu_baseline"]["value"])
PY
bash profiles/r2_final_profile.sh
