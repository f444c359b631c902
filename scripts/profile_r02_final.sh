#!/bin/bash
# Round-2 evidence (run under gpurun, 1 GPU).  Usage: scripts/profile_r02_final.sh r02f
#  1) launch list of the default bench command (every kernel, device time; cold-cache + serialised: compare SHARES)
#  2) ncu --set full of the dominant kernel of each bench line: headline (Florinsky 4 planes, 32768^2), the two kernels of
#     config 4 (9 surface planes; 3x3 windowed planes), the Nuth-Kaab full passes, the variogram pair kernel
#  3) the un-profiled bench lines of both arms
TAG=${1:-r02f}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/launches_${TAG}.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:florinsky_sliding -s 2 -c 1 -o gpurun_out/prof_${TAG}_florinsky4_32768 -f \
    python scripts/prof_one.py 32768 Florinsky slope,aspect,hillshade,curvature 3 > gpurun_out/prof_${TAG}.log 2>&1
ALLC=slope,aspect,hillshade,profile_curvature,tangential_curvature,planform_curvature,flowline_curvature,max_curvature,min_curvature
W4=topographic_position_index,terrain_ruggedness_index,roughness,rugosity
$NCU -k regex:florinsky_sliding -s 2 -c 1 -o gpurun_out/prof_${TAG}_fl9 -f python scripts/prof_one.py 16384 Florinsky $ALLC 3 >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:window3 -s 2 -c 1 -o gpurun_out/prof_${TAG}_win4 -f python scripts/prof_one.py 16384 Florinsky "" 3 $W4 >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:"nkf_dh_kernel|nkf_y_kernel|nk_prepare" -c 5 -o gpurun_out/prof_${TAG}_nkf -f python scripts/nk_fast_prof.py >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:variogram_pairs -c 1 -o gpurun_out/prof_${TAG}_variogram -f \
    python bench_extra.py variogram --n 200000 --cpu-n 2000 --steps 1 >> gpurun_out/prof_${TAG}.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python scripts/perf_probe.py 16384 > gpurun_out/perf_probe_${TAG}.txt 2>&1
python scripts/ncu_brief.py gpurun_out/prof_${TAG}_*.ncu-rep > gpurun_out/ncu_${TAG}.txt 2>&1
tail -c 400 gpurun_out/bench_${TAG}.json
