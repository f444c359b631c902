#!/bin/bash
# Round profiling recipe (run under gpurun, 1 GPU): launch list + full ncu capture of the headline workload + bench lines.
# Usage: scripts/profile_round.sh r01
TAG=${1:-r01}
mkdir -p gpurun_out
# 1) every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_${TAG}.log 2>&1
# 2) the dominant kernel, full set, same workload (32768^2 Florinsky 4 attributes)
ncu --set full --clock-control none --import-source on -k "regex:terrain_fused|florinsky_sliding" -s 3 -c 1 -o gpurun_out/prof_${TAG}_florinsky4_32768 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:terrain_fused -s 3 -c 1 -o gpurun_out/prof_${TAG}_zt4_32768 \
    python bench.py --fit ZevenbergThorne --steps 2 --warmup 3 --no-e2e --no-cpu >> gpurun_out/prof_${TAG}.log 2>&1
# 3) un-profiled bench lines (both arms), all three fits
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --fit ZevenbergThorne --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_zt.json 2>> gpurun_out/bench_${TAG}.err
python bench.py --fit Horn --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${TAG}_horn.json 2>> gpurun_out/bench_${TAG}.err
python scripts/perf_probe.py 16384 > gpurun_out/perf_probe_${TAG}.txt 2>&1
python scripts/perf_vario_nk.py > gpurun_out/perf_vario_nk_${TAG}.txt 2>&1
python bench_extra.py variogram > gpurun_out/bench_${TAG}_variogram.json 2>> gpurun_out/bench_${TAG}.err
python bench_extra.py nuthkaab > gpurun_out/bench_${TAG}_nuthkaab.json 2>> gpurun_out/bench_${TAG}.err
# K2 / K3 kernels under ncu (small workloads; shares + DRAM/issue figures)
ncu --set full --clock-control none -k regex:variogram_pairs -c 1 -o gpurun_out/prof_${TAG}_variogram \
    python bench_extra.py variogram --n 200000 --cpu-n 2000 --steps 1 > gpurun_out/prof_${TAG}_k2.log 2>&1
ncu --set full --clock-control none -k regex:nk_ -s 8 -c 12 -o gpurun_out/prof_${TAG}_nuthkaab \
    python bench_extra.py nuthkaab --size 8192 --cpu-size 256 --steps 1 > gpurun_out/prof_${TAG}_k3.log 2>&1
tail -c 600 gpurun_out/bench_${TAG}.json
