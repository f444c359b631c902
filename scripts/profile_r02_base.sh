#!/bin/bash
# Baseline ncu captures (round 2 start) of the kernels below the roofline: 9-attribute Florinsky (ALG), 3x3 windowed
# indexes with rugosity, 5x5 windowed, Florinsky slope-only (generic kernel).  8192^2 rasters keep the replays short.
TAG=${1:-r02base}
mkdir -p gpurun_out
ALLC=slope,aspect,hillshade,profile_curvature,tangential_curvature,planform_curvature,flowline_curvature,max_curvature,min_curvature
W4=topographic_position_index,terrain_ruggedness_index,roughness,rugosity
W3=topographic_position_index,terrain_ruggedness_index,roughness
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1"
$NCU -k regex:florinsky_sliding -o gpurun_out/prof_${TAG}_fl9 python scripts/prof_one.py 8192 Florinsky $ALLC 3 > gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:terrain_fused -o gpurun_out/prof_${TAG}_win4 python scripts/prof_one.py 8192 Florinsky "" 3 $W4 >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:terrain_fused -o gpurun_out/prof_${TAG}_win3 python scripts/prof_one.py 8192 Florinsky "" 3 $W3 >> gpurun_out/prof_${TAG}.log 2>&1
$NCU -k regex:terrain_fused -o gpurun_out/prof_${TAG}_flslope python scripts/prof_one.py 8192 Florinsky slope 3 >> gpurun_out/prof_${TAG}.log 2>&1
python scripts/perf_probe.py 16384 > gpurun_out/perf_probe_${TAG}.txt 2>&1
tail -25 gpurun_out/perf_probe_${TAG}.txt
