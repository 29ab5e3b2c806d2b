#!/bin/bash
# N-GPU scaling check: bench.py under torchrun at N = $1 (and the NCCL sharding test)
N=${1:-8}
O=gpurun_out
mkdir -p $O
python -m ddsp_b200.build > $O/r02_n${N}_build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv > $O/r02_n${N}_gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N > $O/r02_bench_n${N}.json 2> $O/r02_bench_n${N}.err; echo "bench n$N rc=$?"
cat $O/r02_bench_n${N}.json; tail -3 $O/r02_bench_n${N}.err
