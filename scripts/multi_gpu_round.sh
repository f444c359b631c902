#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus N`): correctness check + weak-scaling terrain bench + sharded NK / variogram.
# Usage: scripts/multi_gpu_round.sh <N> <tag>
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python bench_extra.py nuthkaab --size 4096 --cpu-size 256 --steps 1 > gpurun_out/sanity_nk_${TAG}.json 2> gpurun_out/multi_${TAG}_n${N}.err
$TR --master-port 29511 tests/dist_check_gpu.py > gpurun_out/dist_check_${TAG}_n${N}.log 2>&1
tail -1 gpurun_out/dist_check_${TAG}_n${N}.log
$TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2>> gpurun_out/multi_${TAG}_n${N}.err
$TR --master-port 29513 bench.py --gpus $N --fit ZevenbergThorne --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_${TAG}_zt_n${N}.json 2>> gpurun_out/multi_${TAG}_n${N}.err
$TR --master-port 29514 bench_extra.py nuthkaab > gpurun_out/bench_${TAG}_nuthkaab_n${N}.json 2>> gpurun_out/multi_${TAG}_n${N}.err
$TR --master-port 29515 bench_extra.py variogram > gpurun_out/bench_${TAG}_variogram_n${N}.json 2>> gpurun_out/multi_${TAG}_n${N}.err
for f in gpurun_out/bench_${TAG}_n${N}.json gpurun_out/bench_${TAG}_zt_n${N}.json gpurun_out/bench_${TAG}_nuthkaab_n${N}.json gpurun_out/bench_${TAG}_variogram_n${N}.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); print(sys.argv[1].split("/")[-1], round(d["value"],1), d["unit"], "n_gpus", d["n_gpus"], "ms", round(d["ms_per_step"],2))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
