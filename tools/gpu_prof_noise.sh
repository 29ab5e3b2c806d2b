#!/bin/bash
O=gpurun_out; mkdir -p $O
TAG=${1:-ring}
python tools/numa_probe.py > $O/numa.log 2>&1; cat $O/numa.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'noise_' \
  --launch-skip 2 -c 1 -f -o $O/${TAG}_full python tools/prof_run.py 256 3 > $O/${TAG}_ncu.log 2>&1
tail -3 $O/${TAG}_ncu.log
