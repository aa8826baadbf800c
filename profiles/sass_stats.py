#!/usr/bin/env python
"""Static SASS statistics of the built library: per kernel the instruction count, registers, and how many of the
instructions are packed fp32 (FFMA2 / FADD2 / FMUL2), scalar fp32, MOV, asynchronous copies (LDGSTS = cp.async,
UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA), barriers, and local-memory (spill) accesses.
    python profiles/sass_stats.py [library.so] [substring filter] > profiles/r2/sass_stats.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "spim_registration_b200/libConvolution3D_fftCUDAlib.so"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+).*?SHARED:(\d+) LOCAL:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(4)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
p = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True)
cur, stats = None, {}
for line in p.stdout:
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        stats[cur][m.group(1)] += 1
cols = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MOV", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "UTMALDG", "BAR", "LDL", "STL"]
print(f"{'kernel':70s} {'regs':>4s} {'local':>5s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in cols))
for k, c in sorted(stats.items(), key=lambda kv: demangle(kv[0])):
    name = demangle(k)
    name = re.sub(r"void spim::rt::kernel_entry(_capped)?<spim::", r"\1<", name).replace("(spim::", "(")
    if flt and flt not in name:
        continue
    r = regs.get(k, (0, 0))
    tot = sum(c.values())
    mov = c["MOV"] + sum(v for n, v in c.items() if n.startswith("IMAD") and False)
    print(f"{name[:70]:70s} {r[0]:4d} {r[1]:5d} {tot:6d} " + " ".join(f"{(mov if x == 'MOV' else c[x]):7d}" for x in cols))
