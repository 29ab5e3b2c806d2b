#!/bin/bash
# Final visit of round 2: full GPU tests, smoke(), bench (ours + reference arm + c4), ncu launch list of the
# bench command, ncu --set full of the decoder kernels at B=256, compute-sanitizer memcheck + synccheck.
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
python -m ddsp_b200.build > $O/${TAG}_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --config c4 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --extra 0 > $O/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v3|noise_ring' \
  --launch-skip 4 -c 2 -f -o $O/${TAG}_full_b256 python tools/prof_run.py 256 3 > $O/${TAG}_ncu_full.log 2>&1
ncu -i $O/${TAG}_full_b256.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/${TAG}_memcheck.log 2>&1; tail -3 $O/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py > $O/${TAG}_synccheck.log 2>&1; tail -3 $O/${TAG}_synccheck.log
timeout 300 python tools/reverb_time.py > $O/${TAG}_reverb.log 2>&1; cat $O/${TAG}_reverb.log
timeout 300 python tools/sinusoidal_time.py > $O/${TAG}_sinusoidal.log 2>&1; cat $O/${TAG}_sinusoidal.log
tail -5 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_bench_n1.json; cat $O/${TAG}_bench_reference.json; cat $O/${TAG}_bench_c4.json
