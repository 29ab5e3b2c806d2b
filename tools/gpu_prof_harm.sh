#!/bin/bash
# ncu --set full of the harmonic kernel (B=256): one launch of the decoder path.
O=gpurun_out; mkdir -p $O
TAG=${1:-hv2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harmonic_' \
  --launch-skip 2 -c 1 -f -o $O/${TAG}_full python tools/prof_run.py 256 3 > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
