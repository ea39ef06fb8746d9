#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
V=${1:-ap}
for rep in 1 2; do
for v in libmd2loss.so libmd2loss_$V.so; do
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono
done; done 2>&1 | grep -v Warning | tee gpurun_out/p_times.log
for v in libmd2loss.so libmd2loss_$V.so; do
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 hires
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono structured
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono iid nograd
done 2>&1 | grep -v Warning | tee -a gpurun_out/p_times.log
export MD2_LIB_PATH=$L/libmd2loss_$V.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_decision_locked.py -q -x -m gpu > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log | cut -c1-300
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_p_march $CMD > gpurun_out/p_ncu.log 2>&1
tail -2 gpurun_out/p_ncu.log
