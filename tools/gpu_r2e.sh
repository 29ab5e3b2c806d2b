#!/bin/bash
TAG=${1:-r02e}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
DDSP_B200_HARM_IMPL=v3 timeout 300 python tools/harm_sweep.py > $O/${TAG}_sweep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'harmonic_v|noise_ring' --launch-skip 2 -c 2 --csv --log-file $O/${TAG}_metrics_b256.csv python tools/prof_run.py 256 3 > $O/${TAG}_prof256.log 2>&1
timeout 300 python tools/reverb_time.py > $O/${TAG}_reverb.log 2>&1
tail -6 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_sweep.log; tail -6 $O/${TAG}_metrics_b256.csv | cut -c1-400; cat $O/${TAG}_reverb.log
