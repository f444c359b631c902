#!/bin/bash
# r02a: 3x3 sliding windowed kernel, lean curvature algebra, packed Florinsky masks (1,2,3,4,7,8, all-9/10).
TAG=${1:-r02a}
mkdir -p gpurun_out
python -m pytest tests/test_terrain_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/pytest_${TAG}.txt 2>&1
tail -15 gpurun_out/pytest_${TAG}.txt
python scripts/perf_probe.py 16384 > gpurun_out/perf_probe_${TAG}.txt 2>&1
cat gpurun_out/perf_probe_${TAG}.txt
ALLC=slope,aspect,hillshade,profile_curvature,tangential_curvature,planform_curvature,flowline_curvature,max_curvature,min_curvature
W4=topographic_position_index,terrain_ruggedness_index,roughness,rugosity
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1"
$NCU -k regex:florinsky_sliding -o gpurun_out/prof_${TAG}_fl9 python scripts/prof_one.py 8192 Florinsky $ALLC 3 > gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:window3 -o gpurun_out/prof_${TAG}_win4 python scripts/prof_one.py 8192 Florinsky "" 3 $W4 >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:florinsky_sliding -o gpurun_out/prof_${TAG}_flslope python scripts/prof_one.py 8192 Florinsky slope 3 >> gpurun_out/prof_${TAG}.log 2>&1
