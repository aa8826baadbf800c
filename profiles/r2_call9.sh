timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c9.txt 2>&1; tail -3 gpurun_out/r2_gputests_c9.txt
AB_C3_CFGS="SPIM_NOP=2|SPIM_DEDUP=0" bash profiles/r2_ab.sh r2_ab_c9 "SPIM_NOP=2" "SPIM_DEDUP=0" "SPIM_NOP=2" "SPIM_DEDUP=0 SPIM_CONST_SHIFT=0"
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_c9.jsonl'):
    d=json.loads(l); r=d['r']
    print(d['name'], 'ms_per_step', round(r.get('ms_per_step',0),3), 'value', round(r.get('value',0)/1e9,2), 'checksum', r.get('psi_checksum'))
PY
