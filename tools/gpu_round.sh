#!/bin/bash
# One GPU-box visit: tests, bench lines, ncu launch lists, one full capture.
# usage (from repo root, via gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv \
  --log-file $O/${TAG}_launches_b256.csv python tools/prof_run.py 256 3 > $O/${TAG}_prof256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v2|noise_ring' \
  --launch-skip 4 -c 2 -f -o $O/${TAG}_full_b256 python tools/prof_run.py 256 3 > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n1.json; cat $O/${TAG}_bench_reference.json; cat $O/${TAG}_prof256.log | tail -2
