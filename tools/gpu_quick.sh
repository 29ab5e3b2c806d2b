#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/q_pytest.log
python tools/harm_sweep.py > $O/q_sweep.log 2>&1; cat $O/q_sweep.log
