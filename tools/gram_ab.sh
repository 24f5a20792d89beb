#!/bin/bash
# EuclGram / SC tile kernel timings (and, with arguments, nvcc flag variants): bash tools/gram_ab.sh ["-DGRAM_EUCL_NB=1" ...]
run() { for d in 256 4096; do timeout 120 python tools/bench_metric.py --metric Eucl --n 20000 --dim $d --reps 2 | tail -1 | sed "s/^/$1 /"; timeout 120 python tools/bench_metric.py --metric SC --n 20000 --dim $d --reps 2 | tail -1 | sed "s/^/$1 /"; done; timeout 120 python tools/bench_metric.py --metric Eucl --n 50000 --dim 4096 --reps 2 | sed "s/^/$1 /"; }
run "default"
for v in "$@"; do
  touch phyloligo_b200/csrc/po_gram.cu; PO_NVCC_EXTRA="$v" python phyloligo_b200/build.py > /dev/null 2>&1 && run "$v"
done
