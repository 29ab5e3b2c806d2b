#!/bin/bash
O=gpurun_out; mkdir -p $O
V=tools/variants
timeout 120 python tools/noise_timing.py $V/lib_TAB.so 256 > $O/r3b_timing.log 2>&1
cat $O/r3b_timing.log
timeout 200 python tools/variant_time.py 256 $V/lib_base.so $V/lib_AB.so $V/lib_Bonly.so $V/lib_AB72.so $V/lib_base.so $V/lib_AB.so > $O/r3b_time.log 2>&1
timeout 100 python tools/variant_time.py 32 $V/lib_base.so $V/lib_AB.so $V/lib_AB72.so >> $O/r3b_time.log 2>&1
cat $O/r3b_time.log
cp $V/lib_AB.so ddsp_b200/libddsp_b200.so
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "noise or decoder" > $O/r3b_pytest.log 2>&1; tail -3 $O/r3b_pytest.log
