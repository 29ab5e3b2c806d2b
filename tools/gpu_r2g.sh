#!/bin/bash
TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --config c4 --no-cpu-baseline > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $O/${TAG}_launches_c4.csv python tools/c4_run.py 128 > $O/${TAG}_c4.log 2>&1
tail -8 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_c4.json; tail -3 $O/${TAG}_bench_c4.err; tail -2 $O/${TAG}_c4.log
