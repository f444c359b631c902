#!/bin/bash
# per-kernel device times of one Nuth-Kaab fit at 16384^2 (second fit of nk_breakdown.py: preparation + iterations)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"nk_|nkf_|sel_" -s 372 -c 76 --csv --log-file gpurun_out/nkf_launches.csv python scripts/nk_breakdown.py > /dev/null 2>&1
python - <<"PY"
import csv
rows = list(csv.reader(open("gpurun_out/nkf_launches.csv")))
hdr = None; agg = {}; order = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    key = (d["ID"], d["Kernel Name"][:40])
    agg.setdefault(key, {})[d["Metric Name"]] = (d["Metric Value"], d["Metric Unit"])
    if key not in order: order.append(key)
for k in order:
    m = agg[k]
    t = float(m["gpu__time_duration.sum"][0].replace(",", "")); u = m["gpu__time_duration.sum"][1]
    t_us = t / 1000 if u in ("ns", "nsecond") else t
    print(k[0], k[1], f"{t_us:9.1f} us", m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"))
PY
