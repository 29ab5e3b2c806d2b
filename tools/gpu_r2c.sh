#!/bin/bash
# two-GPU visit: full GPU tests (incl. the NCCL sharding test), bench at N=2 under torchrun, N=1
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
python -m ddsp_b200.build > $O/${TAG}_build.log 2>&1     # rebuilds only if the sources are newer than the .so
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 300 python tools/reverb_time.py > $O/${TAG}_reverb.log 2>&1; cat $O/${TAG}_reverb.log
tail -8 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n2.json; tail -3 $O/${TAG}_bench_n2.err; cat $O/${TAG}_bench_n1.json; cat $O/${TAG}_bench_reference.json
