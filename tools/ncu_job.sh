#!/bin/bash
# ncu --set full captures of the three main kernels (run on the GPU box; each command first runs plainly)
set -x
P1="python tools/bench_profile.py --n 100000 --len 20000 --pattern 1111 --strand both --reps 1"
$P1 > gpurun_out/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:profile_seg -s 1 -c 1 -f -o gpurun_out/r02b_prof_seg_k4 $P1 > gpurun_out/ncu1.log 2>&1
P2="python bench.py --steps 1 --warmup 3 --no-cli --no-extra --no-cpu-baseline"
$P2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jsd_tile -s 3 -c 1 -f -o gpurun_out/r02b_jsd_tile_full $P2 > gpurun_out/ncu2.log 2>&1
P3="python tools/bench_metric.py --metric Eucl --n 20000 --dim 4096 --reps 1"
$P3 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gram_tile -s 1 -c 1 -f -o gpurun_out/r02b_gram_eucl $P3 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
