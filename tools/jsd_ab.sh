#!/bin/bash
# A/B of the JSD chunk variants on the GPU box: default build vs every chunk forced to the two-phase loop.
cd "$(dirname "$0")/.."
for v in default 2 1 0; do
  touch phyloligo_b200/csrc/po_jsd.cu
  if [ "$v" = default ]; then extra=""; else extra="-DJSD_FORCE_VARIANT=$v"; fi
  PO_NVCC_EXTRA="$extra" python -c "from phyloligo_b200 import build; build.build_library()"
  echo "=== variant: $v ($extra) ==="
  python bench.py --scale ${1:-0.3} --steps 5 --warmup 3 --no-cpu-baseline --no-cli --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pairs/s %.4e  ms/step %.2f  roofline frac %.3f  jsd ms %.2f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['stages']['distance_ms_per_step']))"
done
touch phyloligo_b200/csrc/po_jsd.cu
python -c "from phyloligo_b200 import build; build.build_library()"
