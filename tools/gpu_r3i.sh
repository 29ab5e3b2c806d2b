#!/bin/bash
# last visit: the bench line with the worst-case extra, smoke(), the GPU suite at HEAD
O=gpurun_out; mkdir -p $O
timeout 150 python bench.py > $O/r3i_bench_n1.json 2> $O/r3i_bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/r3i_bench_n1.json
timeout 50 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3i_smoke.log 2>&1; tail -n 1 $O/r3i_smoke.log
timeout 100 python -m pytest tests -m gpu -q -x --timeout 60 > $O/r3i_pytest.log 2>&1; tail -n 2 $O/r3i_pytest.log
