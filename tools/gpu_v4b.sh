#!/bin/bash
O=gpurun_out; mkdir -p $O
./tools/microbench3 > $O/v4b_microbench3.txt 2>&1
cat $O/v4b_microbench3.txt
timeout 600 python tools/variant_time.py 256 tools/variants/lib_V3.so tools/variants/lib_A_f64_oeoe.so tools/variants/lib_C_fixed.so tools/variants/lib_D_ctas5.so tools/variants/lib_E_ctas7.so tools/variants/lib_F_ctas8.so tools/variants/lib_V3.so > $O/v4b_time.log 2>&1
cat $O/v4b_time.log
