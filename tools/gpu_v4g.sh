#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/v4g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v4g_pytest.log
tail -3 $O/v4g_pytest.log
for fw in 7 8 11 15; do
  DDSP_B200_HARM_FW=$fw timeout 300 python tools/variant_time.py 256 tools/variants/lib_I_nw1.so tools/variants/lib_I_nw1r.so tools/variants/lib_I_nw2.so tools/variants/lib_I_nw4.so 2>&1 | sed "s/^/FW=$fw /" >> $O/v4g_time.log
done
cat $O/v4g_time.log
