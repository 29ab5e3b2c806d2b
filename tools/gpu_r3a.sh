#!/bin/bash
# diagnostic visit: noise_ring phase counters, producer-group / slot variants, one full capture of both decoder kernels
O=gpurun_out; mkdir -p $O
V=tools/variants
timeout 120 python tools/noise_timing.py $V/lib_T.so 256 > $O/r3a_timing.log 2>&1
timeout 120 python tools/noise_timing.py $V/lib_TG2S8.so 256 >> $O/r3a_timing.log 2>&1
cat $O/r3a_timing.log
timeout 200 python tools/variant_time.py 256 $V/lib_base.so $V/lib_G2S7.so $V/lib_G2S8.so $V/lib_G2S8r.so $V/lib_base.so > $O/r3a_time.log 2>&1
timeout 100 python tools/variant_time.py 32 $V/lib_base.so $V/lib_G2S8.so $V/lib_G2S8r.so >> $O/r3a_time.log 2>&1
cat $O/r3a_time.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v4|noise_ring' --launch-skip 4 -c 2 -f -o $O/r3a_full_b256 python tools/prof_run.py 256 3 > $O/r3a_ncu.log 2>&1; echo "full rc=$?"
ls -la $O | tail
