#!/bin/bash
# Stage the UNMODIFIED reference (read-only at /root/reference in the build container) under baseline/_ref/ so that it travels
# to the GPU box with gpurun (baseline/_ref is git-ignored: reference sources never enter this repo's history).  Used there by
# oracle/ref_shims.py for (i) the reference-GPU-eager timing of bench.py (the 10x denominator of BASELINE.json's north_star),
# (ii) the drop-in test that runs the reference's own train_one_epoch on the CUDA model, (iii) large-size parity checks.
set -e
SRC="${1:-/root/reference}"
DST="$(dirname "$0")/../baseline/_ref"
[ -d "$SRC/src/models" ] || { echo "no reference tree at $SRC"; exit 1; }
mkdir -p "$DST"
rm -rf "$DST/src" "$DST/evaluation"
cp -r "$SRC/src" "$DST/src"
cp -r "$SRC/evaluation" "$DST/evaluation"
find "$DST" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
echo "staged $(du -sh "$DST" | cut -f1) at $DST"
