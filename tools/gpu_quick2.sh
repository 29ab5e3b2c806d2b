#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/q_pytest.log
{
DDSP_B200_NOISE_IMPL=pipe python tools/harm_sweep.py
python tools/harm_sweep.py
} > $O/q_sweep.log 2>&1
cat $O/q_sweep.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/q_memcheck.log 2>&1; tail -4 $O/q_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/q_racecheck.log 2>&1; tail -4 $O/q_racecheck.log
