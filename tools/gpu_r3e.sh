#!/bin/bash
O=gpurun_out; mkdir -p $O
V=tools/variants
timeout 120 python tools/variant_time.py 256 $V/lib_base.so $V/lib_Q.so $V/lib_Q168.so $V/lib_base.so $V/lib_Q.so > $O/r3e_time.log 2>&1
timeout 60 python tools/variant_time.py 32 $V/lib_base.so $V/lib_Q.so >> $O/r3e_time.log 2>&1
cat $O/r3e_time.log
timeout 60 python tools/noise_timing.py $V/lib_TQ.so 256 > $O/r3e_timing.log 2>&1
cat $O/r3e_timing.log
cp $V/lib_Q.so ddsp_b200/libddsp_b200.so
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 90 > $O/r3e_pytest_Q.log 2>&1; tail -n 3 $O/r3e_pytest_Q.log
