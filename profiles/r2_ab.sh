#!/bin/bash
# A/B timing of kernel switches with the torch-free variant child of bench.py (per-kernel CUDA-event times, 44*Np fraction):
#   bash profiles/r2_ab.sh <tag> "<ENV=V ...>" "<ENV=V ...>" ...      -> gpurun_out/<tag>.jsonl (one line per configuration)
set -u
tag=$1; shift
out=gpurun_out/$tag.jsonl
: > $out
run() {  # name, size args, env...
    local name=$1; shift
    local size=$1; shift
    local line
    line=$(env "$@" timeout 200 python bench.py --variant-child $size 2>gpurun_out/$tag.err | grep '^{' | tail -1)
    [ -z "$line" ] && line="{\"error\": \"$(tail -c 300 gpurun_out/$tag.err | tr '\n"' '  ')\"}"
    echo "{\"name\": \"$name\", \"env\": \"$*\", \"r\": $line}" >> $out
    python - "$name" "$line" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(sys.argv[1], "frac", round(d.get("conv_pass_frac") or 0, 4), {k[:6]: round(v, 4) for k, v in (d.get("per_kernel_ms") or {}).items()}, {k[10:]: round(v, 4) for k, v in (d.get("x_inv_by_epilogue_ms") or {}).items()}, d.get("psi_checksum"), d.get("error"))
PY
}
C1="--views 7 --brick 256 512 512"
C3="--views 6 --brick 512 1024 1024 --iter-type 0 --variant-iters 2"
for cfg in "$@"; do
    run "c1/$cfg" "$C1" $cfg SPIM_NOP=1
done
# configs[2] volume on one GPU: AB_C3_CFGS="cfg|cfg|..."
if [ -n "${AB_C3_CFGS:-}" ]; then
    IFS='|' read -ra c3cfgs <<< "$AB_C3_CFGS"
    for cfg in "${c3cfgs[@]}"; do
        run "c3/$cfg" "$C3" $cfg SPIM_NOP=1
    done
fi
