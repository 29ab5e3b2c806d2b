#!/bin/bash
# Validation visit: bench lines first (decoder, configs[3]), ncu launch list of the bench command, one full
# capture of the two decoder kernels, noise_ring phase counters, then the whole GPU test suite.
O=gpurun_out; mkdir -p $O; T=r3g
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${T}_gpu.txt 2>&1
timeout 240 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/${T}_bench_n1.json
timeout 150 python bench.py --config c4 > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err; echo "c4 rc=$?"; cut -c1-300 $O/${T}_bench_c4.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra 0 > $O/${T}_bench_ncu.log 2>&1; echo "launch list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v4|noise_ring' --launch-skip 4 -c 2 -f -o $O/${T}_full_b256 python tools/prof_run.py 256 3 > $O/${T}_ncu.log 2>&1; echo "full rc=$?"
timeout 60 python tools/noise_timing.py tools/variants/lib_T.so 256 > $O/${T}_timing.log 2>&1; cat $O/${T}_timing.log
timeout 500 python -m pytest tests -m gpu -q -x --timeout 90 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 4 $O/${T}_pytest.log
timeout 90 python tools/variant_time.py 256 tools/variants/lib_final.so tools/variants/lib_R176.so tools/variants/lib_R160.so tools/variants/lib_final.so > $O/${T}_regsplit.log 2>&1; cat $O/${T}_regsplit.log
