#!/bin/bash
# Rebuild po_gram.cu with different ring geometries and time EuclGram / SC (run on the GPU box).
# usage: tools/gram_variants.sh "KS:STAGES ..."   (96 KB of ring per CTA keeps two CTAs per SM)
set -e
cd "$(dirname "$0")/.."
for v in ${1:-"32:3 16:6 32:6"}; do
  ks=${v%%:*}; st=${v##*:}
  echo "=== GRAM_EUCL_KS=$ks GRAM_EUCL_STAGES=$st ==="
  touch phyloligo_b200/csrc/po_gram.cu
  PO_NVCC_EXTRA="-DGRAM_EUCL_KS=$ks -DGRAM_EUCL_STAGES=$st" python -c "from phyloligo_b200 import build; build.build_library()"
  python tools/bench_metric.py --metric Eucl --n 50000 --dim 4096 --reps 3
  python tools/bench_metric.py --metric Eucl --n 20000 --dim 4096 --reps 3
  python tools/bench_metric.py --metric Eucl --n 20000 --dim 256 --reps 3
done
touch phyloligo_b200/csrc/po_gram.cu
python -c "from phyloligo_b200 import build; build.build_library()"
python tools/bench_metric.py --metric SC --n 20000 --dim 4096 --reps 3
python tools/bench_metric.py --metric SC --n 20000 --dim 256 --reps 3
