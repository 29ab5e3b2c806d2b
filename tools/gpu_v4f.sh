#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v4|noise_ring' \
  --launch-skip 4 -c 2 -f -o $O/v4f_full_b256 python tools/prof_run.py 256 3 > $O/v4f_ncu.log 2>&1
tail -3 $O/v4f_ncu.log
ls -la $O/
