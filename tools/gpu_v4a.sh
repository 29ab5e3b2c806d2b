#!/bin/bash
# harmonic_v4 first visit: parity tests, then v4 / v3 timings side by side.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/v4a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v4a_pytest.log
tail -15 $O/v4a_pytest.log
TAG=v4 timeout 300 python tools/harm_time.py > $O/v4a_time.log 2>&1
DDSP_B200_HARM_IMPL=v3 TAG=v3 timeout 300 python tools/harm_time.py >> $O/v4a_time.log 2>&1
cat $O/v4a_time.log
