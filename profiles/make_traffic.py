#!/usr/bin/env python
"""profiles/traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full --page raw --csv` dump of one
view-step of the hot path (profiles/r2_ncu.sh): dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by sweep.
    python profiles/make_traffic.py profiles/r2/r2f_default_raw.csv.gz > profiles/traffic.json"""
import csv
import gzip
import io
import json
import sys

path = sys.argv[1]
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


# launch order of one view-step: x-fwd, y-fwd, z-mid, y-inv, x-inv(ratio), x-fwd, y-fwd, z-mid, y-inv, x-inv(update)
names = ["x_fwd_r2c", "y_fwd", "z_fwd_mul_inv", "y_inv", "x_inv_ratio"] + ["x_fwd_r2c", "y_fwd", "z_fwd_mul_inv", "y_inv", "x_inv_update"]
assert len(data) == 10, f"expected the 10 launches of one view-step, found {len(data)}"
per, kern = {}, {}
for r, nm in zip(data, names):
    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    per.setdefault(nm, []).append(b)
    kern.setdefault(nm, []).append(r[col["Kernel Name"]])
avg = {k: sum(v) / len(v) for k, v in per.items()}
out = {"x_fwd_r2c": avg["x_fwd_r2c"], "y_fwd": avg["y_fwd"], "z_fwd_mul_inv": avg["z_fwd_mul_inv"], "y_inv": avg["y_inv"],
       "x_inv_c2r_epilogue": (avg["x_inv_ratio"] + avg["x_inv_update"]) / 2,
       "_detail": {"x_inv_ratio": avg["x_inv_ratio"], "x_inv_update": avg["x_inv_update"],
                   "conv1": {n: per[n][0] for n in ("x_fwd_r2c", "y_fwd", "z_fwd_mul_inv", "y_inv")},
                   "conv2": {n: per[n][1] for n in ("x_fwd_r2c", "y_fwd", "z_fwd_mul_inv", "y_inv")}},
       "_kernels": {k: sorted(set(v)) for k, v in kern.items()},
       "_source": f"{path}: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, "
                  "one view-step of the 512x512x256 brick, 31^3 PSF"}
print(json.dumps(out, indent=1))
