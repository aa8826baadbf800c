timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c10.txt 2>&1; tail -3 gpurun_out/r2_gputests_c10.txt
AB_C3_CFGS="SPIM_NOP=2|SPIM_XFWD_T1=512|SPIM_XFWD_T1=640|SPIM_COLP=2" bash profiles/r2_ab.sh r2_ab_c10 "SPIM_NOP=2"
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_c10.jsonl'):
    d=json.loads(l); r=d['r']
    print(d['name'], 'ms_per_step', round(r.get('ms_per_step',0),3), 'value', round(r.get('value',0)/1e9,2), 'checksum', r.get('psi_checksum'))
PY
