#!/bin/bash
# Short K1-only evidence run (launch list + full ncu capture of the headline kernel + the default bench line).
# Usage: scripts/profile_quick.sh r01d
TAG=${1:-r01d}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:terrain_fused|florinsky_sliding" -s 3 -c 1 -o gpurun_out/prof_${TAG}_florinsky4_32768 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_${TAG}.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench_extra.py binning --steps 2 > gpurun_out/bench_${TAG}_binning.json 2>> gpurun_out/bench_${TAG}.err
tail -c 700 gpurun_out/bench_${TAG}.json
