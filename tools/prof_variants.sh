#!/bin/bash
# Build and time profiling-kernel variants on the GPU box (tuning aid).
for v in "-DPO_SEG_LANE_BYTES=96 -DPO_SEG_NSTAGE=1" "-DPO_SEG_LANE_BYTES=96 -DPO_SEG_NSTAGE=2" "-DPO_SEG_LANE_BYTES=128 -DPO_SEG_NSTAGE=1" "-DPO_SEG_LANE_BYTES=168 -DPO_SEG_NSTAGE=1" "-DPO_SEG_LANE_BYTES=256 -DPO_SEG_NSTAGE=1"; do
  touch phyloligo_b200/csrc/po_profile_seg.cu
  PO_NVCC_EXTRA="$v" python phyloligo_b200/build.py > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  python bench.py --scale 0.5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$v', 'prof ms %.4f'%d['stages']['profiling_ms_per_launch'], 'Gbase/s %.0f'%d['stages']['profiling_gbases_per_s_rank0'])"
done
