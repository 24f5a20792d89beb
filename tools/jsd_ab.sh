#!/bin/bash
# A/B of JSD kernel variants on the GPU box: bash tools/jsd_ab.sh "<nvcc flags>" "<nvcc flags>" ...
run() {
  python bench.py --scale 0.3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1', 'C2@30k: %.3e pairs/s'%d['value'], 'dist ms', round(d['stages']['distance_ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))"
  python tools/run_config.py C5 --scale 0.05 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1', 'C5@50k:', {k:v for k,v in d['JSD'].items() if 'kernel' in k or 'pairs_per_s' in k})"
}
run default
for v in "$@"; do
  touch phyloligo_b200/csrc/po_jsd.cu
  PO_NVCC_EXTRA="$v" python phyloligo_b200/build.py > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  run "$v"
done
