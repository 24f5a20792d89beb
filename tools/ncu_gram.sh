#!/bin/bash
set -x
P3="python tools/bench_metric.py --metric Eucl --n 20000 --dim 4096 --reps 1"
$P3 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gram_tile -s 1 -c 1 -f -o gpurun_out/r02c_gram_eucl8 $P3 > gpurun_out/ncu3.log 2>&1
P4="python tools/bench_metric.py --metric SC --n 20000 --dim 4096 --reps 1"
$P4 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gram_tile -s 1 -c 1 -f -o gpurun_out/r02c_gram_sc8 $P4 > gpurun_out/ncu4.log 2>&1
