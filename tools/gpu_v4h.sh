#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/v4h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v4h_pytest.log
tail -4 $O/v4h_pytest.log
for B in 256 32; do
timeout 300 python tools/variant_time.py $B tools/variants/lib_I_nw4.so tools/variants/lib_J_p64.so tools/variants/lib_J_p72.so tools/variants/lib_J_p80c136.so tools/variants/lib_V3.so >> $O/v4h_time.log 2>&1
done
cat $O/v4h_time.log
