#!/bin/bash
TAG=${1:-r02i}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_effects.py tests/test_gpu_host.py -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 300 python tools/reverb_time.py > $O/${TAG}_reverb.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extra 0 > $O/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $O/${TAG}_launches_reverb.csv python tools/reverb_time.py > /dev/null 2>&1
tail -5 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_reverb.log
