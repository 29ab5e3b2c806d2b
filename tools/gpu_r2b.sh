#!/bin/bash
# round-2 kernel iteration visit: full GPU tests, A/B of the harmonic generations,
# instruction counts, one full ncu capture of the decoder kernels at B=256
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
for impl in v3 v2; do
  DDSP_B200_HARM_IMPL=$impl timeout 300 python tools/harm_sweep.py >> $O/${TAG}_sweep.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'harmonic_v|noise_ring' --launch-skip 2 -c 4 --csv --log-file $O/${TAG}_metrics_b256.csv python tools/prof_run.py 256 3 > $O/${TAG}_prof256.log 2>&1
DDSP_B200_HARM_IMPL=v2 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'harmonic_v' --launch-skip 1 -c 2 --csv --log-file $O/${TAG}_metrics_b256_v2.csv python tools/prof_run.py 256 3 >> $O/${TAG}_prof256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v3|noise_ring' \
  --launch-skip 4 -c 2 -f -o $O/${TAG}_full_b256 python tools/prof_run.py 256 3 > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/${TAG}_full_b256.ncu-rep --page source --csv > $O/${TAG}_source.csv 2>/dev/null
ncu -i $O/${TAG}_full_b256.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
tail -15 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_sweep.log; cat $O/${TAG}_metrics_b256.csv | tail -20; cat $O/${TAG}_metrics_b256_v2.csv | tail -8
