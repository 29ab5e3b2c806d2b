#!/bin/bash
# harmonic v3 rework (thread = frame prologue, mask-free "clean" frames, 32-frame tiles):
# parity tests, then timings for the occupancy / tile-size variants, one ncu --set full capture.
TAG=${1:-r02j}
O=gpurun_out
mkdir -p $O
python -m ddsp_b200.build --force > $O/${TAG}_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_backward.py -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
for fw in 0 16 4; do
  if [ $fw = 0 ]; then unset DDSP_B200_HARM_FW; else export DDSP_B200_HARM_FW=$fw; fi
  TAG="ctas6" timeout 300 python tools/harm_time.py >> $O/${TAG}_harm_time.log 2>&1
done
unset DDSP_B200_HARM_FW
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'harmonic_v3' \
  --launch-skip 2 -c 1 -f -o $O/${TAG}_full_b256 python tools/prof_run.py 256 3 > $O/${TAG}_ncu_full.log 2>&1
DDSP_B200_NVCC_EXTRA="-DDDSP_HV3_MIN_CTAS=5" python -m ddsp_b200.build --force >> $O/${TAG}_build.log 2>&1
for fw in 0 16; do
  if [ $fw = 0 ]; then unset DDSP_B200_HARM_FW; else export DDSP_B200_HARM_FW=$fw; fi
  TAG="ctas5" timeout 300 python tools/harm_time.py >> $O/${TAG}_harm_time.log 2>&1
done
unset DDSP_B200_HARM_FW
DDSP_B200_NVCC_EXTRA="-DDDSP_HV3_MIN_CTAS=4" python -m ddsp_b200.build --force >> $O/${TAG}_build.log 2>&1
TAG="ctas4" timeout 300 python tools/harm_time.py >> $O/${TAG}_harm_time.log 2>&1
for ns in 1000 20000; do
  DDSP_B200_NVCC_EXTRA="-DDDSP_MBAR_HINT_NS=$ns" python -m ddsp_b200.build --force >> $O/${TAG}_build.log 2>&1
  TAG="ctas6_hint$ns" timeout 300 python tools/harm_time.py >> $O/${TAG}_harm_time.log 2>&1
done
python -m ddsp_b200.build --force >> $O/${TAG}_build.log 2>&1
cat $O/${TAG}_harm_time.log
