#!/usr/bin/env python
"""profiles/make_traffic.py <launches.csv> <commit> -> profiles/traffic_r02.json

Reads an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary` and sums the DRAM bytes of the kernels of
ONE device-resident cfg2 step (the first full set of mbd_* launches).  bench.py reads the result as
`roofline.traffic` -- a measurement of the current kernels, tagged with the commit it was captured at."""
import csv
import json
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
recs = {}
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    k = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("sd::", ""))
    recs.setdefault(k, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
step, seen = [], set()
for (i, name), m in sorted(recs.items()):
    if not name.startswith("mbd_"):  # torch kernels that generate the synthetic walks
        continue
    if name in seen:  # second step begins
        break
    seen.add(name)
    step.append(dict(kernel=name, us=m["gpu__time_duration.sum"] / 1e3, dram_read_MB=m["dram__bytes_read.sum"] / 1e6,
                     dram_write_MB=m["dram__bytes_write.sum"] / 1e6))
total = sum(k["dram_read_MB"] + k["dram_write_MB"] for k in step) * 1e6
alg = 8.0 * 100_000 * 1024 + 8.0 * 100_000
out = dict(commit=sys.argv[2], source=os.path.basename(sys.argv[1]), workload="cfg2 step, 100 000 curves x 1024 rows, device resident",
           dram_bytes_per_step=total, algorithmic_bytes_per_step=alg, ratio=total / alg,
           kernel_us_sum_under_ncu=sum(k["us"] for k in step), kernels=step)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic_r02.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
