#!/bin/bash
O=gpurun_out; mkdir -p $O
V=tools/variants
timeout 60 python tools/noise_timing.py $V/lib_TABP.so 256 > $O/r3c_timing.log 2>&1
cat $O/r3c_timing.log
timeout 150 python tools/variant_time.py 256 $V/lib_AB.so $V/lib_ABP.so $V/lib_G2S8r.so $V/lib_G2S8rP.so $V/lib_NW8.so $V/lib_AB.so > $O/r3c_time.log 2>&1
cat $O/r3c_time.log
