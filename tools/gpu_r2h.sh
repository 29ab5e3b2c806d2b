#!/bin/bash
TAG=${1:-r02h}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --config c4 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
tail -8 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n1.json; tail -3 $O/${TAG}_bench_n1.err; cat $O/${TAG}_bench_c4.json; tail -3 $O/${TAG}_bench_c4.err
