#!/bin/bash
O=gpurun_out; mkdir -p $O
for fw in 8 4; do
  DDSP_B200_HARM_FW=$fw timeout 300 python tools/variant_time.py 256 ddsp_b200/libddsp_b200.so tools/variants/lib_V3.so 2>&1 | sed "s/^/FW=$fw /" >> $O/v4c_time.log
done
cat $O/v4c_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v4' \
  --launch-skip 2 -c 1 -f -o $O/v4c_full_b256 python tools/prof_run.py 256 3 > $O/v4c_ncu.log 2>&1
tail -3 $O/v4c_ncu.log
ls -la $O/
