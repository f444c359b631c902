#!/bin/bash
# End-of-round evidence on one GPU: full GPU test log, parity report on the final kernels, both bench arms.
TAG=${1:-r02g}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_${TAG}.txt
tail -2 gpurun_out/pytest_gpu_${TAG}.txt
python scripts/parity_report.py > gpurun_out/parity_report_${TAG}.txt 2>&1
tail -3 gpurun_out/parity_report_${TAG}.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - gpurun_out/bench_${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("headline", round(d["value"]), "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), d["clocks"]["reasons"])
e = d["e2e"]
print("e2e", round(e["value"]), "pinned", round(e["pinned"]["value"]), e.get("d2h_link_probe"))
for k, v in d.get("extra", {}).items():
    print(k, round(v.get("value", 0), 1), "ms", round(v.get("ms_per_step", 0), 3), "frac", round((v.get("roofline") or {}).get("frac", 0), 3), v.get("error"), v.get("default_subsample"))
PY
