#!/bin/bash
# Put an UNMODIFIED copy of the reference's Python tree where the GPU box can see it (baseline/_ref is git-ignored, never
# committed, but travels with gpurun snapshots): the drop-in tests (tests/test_gpu_dropin.py) import the reference's own
# model file from there and run it against this repository's `layers`.  The reference has no setup.py / pyproject, so the
# `pip install --target baseline/_ref` of the base contract does not apply; this is a plain copy of ssd_liverdet/.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -d "$SRC/ssd_liverdet" ] || { echo "no reference tree at $SRC"; exit 1; }
rm -rf "$DST"; mkdir -p "$DST"
cp -r "$SRC/ssd_liverdet" "$DST/ssd_liverdet"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "reference copied to $DST ($(du -sh "$DST" | cut -f1))"
