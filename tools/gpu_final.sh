#!/bin/bash
# last visit of the round: sanitizer on the small end-to-end run, smoke(), a short bench
O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/f_memcheck.log 2>&1; tail -3 $O/f_memcheck.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 10 > $O/f_bench.json 2> $O/f_bench.err; echo "bench rc=$?"; cut -c1-400 $O/f_bench.json; tail -2 $O/f_bench.err
