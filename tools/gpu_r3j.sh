#!/bin/bash
# frames per one-warp CTA of the harmonic kernel (runtime knob DDSP_B200_HARM_FW), one process each
O=gpurun_out; mkdir -p $O
for fw in 0 15 7 13 9; do
  echo "DDSP_B200_HARM_FW=$fw" >> $O/r3j_fw.log
  DDSP_B200_HARM_FW=$fw timeout 45 python tools/variant_time.py 256 ddsp_b200/libddsp_b200.so >> $O/r3j_fw.log 2>&1
done
cat $O/r3j_fw.log
