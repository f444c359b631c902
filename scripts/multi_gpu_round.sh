#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus N`): correctness check (sharded == single GPU) + the bench line with the
# strong-scaled extras (c4: one 65536^2 raster, all 13 planes; c5: sharded Nuth-Kaab; c3: split variogram).
# Usage: scripts/multi_gpu_round.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tests/dist_check_gpu.py > gpurun_out/dist_check_${TAG}_n${N}.log 2>&1
tail -2 gpurun_out/dist_check_${TAG}_n${N}.log
$TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/multi_${TAG}_n${N}.err
tail -c 600 gpurun_out/multi_${TAG}_n${N}.err
python - gpurun_out/bench_${TAG}_n${N}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("headline", round(d["value"]), d["unit"], "n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3))
e = d.get("e2e") or {}
print("e2e", round(e.get("value", 0)), "ms", round(e.get("ms_per_step", 0), 1), "d2h GB/s per GPU", round(e.get("d2h_gbs_per_gpu", 0), 1), e.get("host_binding"))
for k, v in d.get("extra", {}).items():
    print(k, round(v.get("value", 0), 1) if "value" in v else None, v.get("unit"), "ms", round(v.get("ms_per_step", 0), 2), "frac", round((v.get("roofline") or {}).get("frac", 0), 3), v.get("error"))
PY
