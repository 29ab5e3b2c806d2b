#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/v4d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v4d_pytest.log
tail -3 $O/v4d_pytest.log
for fw in 16 8 4; do
  DDSP_B200_HARM_FW=$fw timeout 300 python tools/variant_time.py 256 tools/variants/lib_G_base.so tools/variants/lib_G_nw8.so tools/variants/lib_G_ctas5.so tools/variants/lib_G_ctas7.so tools/variants/lib_V3.so 2>&1 | sed "s/^/FW=$fw /" >> $O/v4d_time.log
done
cat $O/v4d_time.log
