#!/bin/bash
# quick GPU visit: parity tests + kernel sweeps (no profiler)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/q_pytest.log
{
DDSP_B200_HARM_IMPL=fast python tools/harm_sweep.py
for fw in 0 2 4 8 16; do DDSP_B200_HARM_FW=$fw python tools/harm_sweep.py; done
python tools/e2e_sweep.py 32
} > $O/q_sweep.log 2>&1
cat $O/q_sweep.log
