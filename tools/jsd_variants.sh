#!/bin/bash
# Build and time JSD kernel variants on the GPU box (tuning aid).
for v in "-DJSD_UNROLL=1" "-DJSD_UNROLL=4" "-DJSD_STAGES=2 -DJSD_CTAS=5" "-DJSD_STAGES=2 -DJSD_CTAS=6" "-DJSD_UNROLL=2"; do
  touch phyloligo_b200/csrc/po_jsd.cu
  PO_NVCC_EXTRA="$v" python phyloligo_b200/build.py > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  python bench.py --scale 0.2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', '%.3e pairs/s'%d['value'], 'dist ms', d['stages']['distance_ms_per_step'])"
done
