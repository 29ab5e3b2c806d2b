#!/bin/bash
# round-2 first visit: tests, new bench (B=256, graph), reference arm quick, launch list
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extra 0 > $O/${TAG}_bench_under_ncu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
