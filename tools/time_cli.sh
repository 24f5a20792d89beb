#!/bin/bash
# Wall-clock of the command line on a synthetic assembly: bash tools/time_cli.sh <contigs> <mean_len> <metric> <large> [extra flags]
set -e
N=${1:-30000}; L=${2:-20000}; M=${3:-JSD}; LARGE=${4:-memmap}; shift 4 || true
D=$(mktemp -d /tmp/po_cli_XXXX)
python - <<PY
import sys; sys.path.insert(0, "."); import numpy as np
from phyloligo_b200 import synth
fasta, total = synth.fast_fasta_bytes($N, $L, seed=2)
np.asarray(fasta).tofile("$D/asm.fasta"); print("fasta bytes", len(fasta), "bases", total)
PY
for rep in 1 2; do
  T0=$(date +%s.%N)
  PO_VERBOSE=1 python -m phyloligo_b200.phyloligo -i $D/asm.fasta -d $M --method joblib --large $LARGE -o $D/out.mat -w $D "$@" 2>&1 | grep -v "^Using\|^Computing\|^Writing"
  python -c "import time,sys; print(\"wall %.2f s\" % (time.time() - float(sys.argv[1])))" $T0; echo " (includes interpreter start and import torch)"
done
ls -la $D/out.mat | awk '{print "output bytes", $5}'
rm -rf $D
