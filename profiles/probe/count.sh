#!/bin/bash
# usage: profiles/probe/count.sh "-DPR=8 -DPINV=0 -DPTW=1 -DPSRC=0 -DPDST=0" [kernel]
# prints the instruction mix of the probe kernel's main loop body (between the loop head label and the backward branch)
set -e
cd "$(dirname "$0")"
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -cubin $1 stage_probe.cu -o /tmp/stage_probe.cubin
cuobjdump -sass /tmp/stage_probe.cubin > /tmp/stage_probe.sass
python3 - "$2" <<'PY'
import re, sys, collections
want = sys.argv[1] or "probe_stage"
txt = open("/tmp/stage_probe.sass").read()
fn = [f for f in txt.split("Function : ")[1:] if f.startswith(want)][0]
ins = []
for ln in fn.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m: ins.append((int(m.group(1), 16), m.group(3), ln))
# main loop = the longest backward branch
best = None
for a, op, ln in ins:
    if op.startswith("BRA"):
        t = re.search(r"0x([0-9a-f]+)", ln.split("BRA")[1])
        if t:
            tgt = int(t.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
body = [i for i in ins if best and best[0] <= i[0] <= best[1]] if best else ins
c = collections.Counter()
for _, op, _ in body:
    b = op.split(".")[0]
    cls = ("FP" if b in ("FADD","FMUL","FFMA","FMNMX","FSEL","FSETP","FCHK","MUFU","FADD2","FMUL2","FFMA2") else
           "LDS/STS" if b in ("LDS","STS","LDSM") else "LDG/STG" if b in ("LDG","STG","LDGSTS","LD","ST") else
           "INT" if b in ("IMAD","IADD3","IADD","LEA","SHF","LOP3","ISETP","IABS","I2F","F2I","IMNMX","SEL","VIADD","PRMT","SGXT","LOP","UIADD3","UIMAD","ULEA","USHF","ULOP3","UISETP","UMOV","USEL") else
           "MOV" if b in ("MOV","R2UR","S2R","S2UR","CS2R","LDC","LDCU","ULDC","R2P","P2R") else
           "CTRL" if b in ("BRA","BSSY","BSYNC","BAR","EXIT","WARPSYNC","NOP","CALL","RET","BREAK","YIELD") else "OTHER:"+b)
    c[cls] += 1
print(f"{want}: loop body {len(body)} instructions (whole kernel {len(ins)})", dict(c))
PY
