#!/bin/bash
# validation visit: full GPU test suite (per-test timeout), bench lines, launch list, one full capture
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q -x --timeout 90 > $O/v4i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v4i_pytest.log
tail -4 $O/v4i_pytest.log
timeout 300 python bench.py > $O/v4i_bench.json 2> $O/v4i_bench.err; echo "bench rc=$?"; cut -c1-600 $O/v4i_bench.json
timeout 200 python bench.py --config c4 > $O/v4i_bench_c4.json 2> $O/v4i_bench_c4.err; echo "c4 rc=$?"; cut -c1-300 $O/v4i_bench_c4.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/v4i_launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/v4i_bench_ncu.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v4|noise_ring' --launch-skip 4 -c 2 -f -o $O/v4i_full_b256 python tools/prof_run.py 256 3 > $O/v4i_ncu.log 2>&1; echo "full rc=$?"
ls -la $O | tail -12
