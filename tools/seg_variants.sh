#!/bin/bash
# staging geometry of the segment profiling kernel (GPU box): bytes per lane per tile, staging buffers per warp
cd "$(dirname "$0")/.."
for v in "168 1" "168 2" "252 1" "336 1" "336 2" "84 2"; do
  set -- $v
  touch phyloligo_b200/csrc/po_profile_seg.cu
  PO_NVCC_EXTRA="-DPO_SEG_LANE_BYTES=$1 -DPO_SEG_NSTAGE=$2" python -c "from phyloligo_b200 import build; build.build_library()"
  echo "=== LANE_BYTES=$1 NSTAGE=$2 ==="
  python tools/bench_profile.py --n 100000 --len 20000 --pattern 1111 --strand both
  python tools/bench_profile.py --n 200000 --len 5000 --pattern 11111 --strand both
done
touch phyloligo_b200/csrc/po_profile_seg.cu
python -c "from phyloligo_b200 import build; build.build_library()"
