#!/bin/bash
O=gpurun_out; mkdir -p $O
V=tools/variants
timeout 120 python tools/variant_time.py 256 $V/lib_AB.so $V/lib_WF.so $V/lib_G2r.so $V/lib_G2rWF.so $V/lib_G2qWF.so $V/lib_AB.so > $O/r3d_time.log 2>&1
timeout 60 python tools/variant_time.py 32 $V/lib_AB.so $V/lib_G2rWF.so $V/lib_G2qWF.so >> $O/r3d_time.log 2>&1
cat $O/r3d_time.log
timeout 60 python tools/noise_timing.py $V/lib_TG2qWF.so 256 > $O/r3d_timing.log 2>&1
cat $O/r3d_timing.log
for L in G2qWF G2rWF; do
cp $V/lib_$L.so ddsp_b200/libddsp_b200.so
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -k "noise or decoder" > $O/r3d_pytest_$L.log 2>&1; tail -2 $O/r3d_pytest_$L.log
done
