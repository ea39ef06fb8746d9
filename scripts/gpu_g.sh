#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/g_pytest.log
tail -3 gpurun_out/g_pytest.log
for v in a2c4 a1c5 a3c3; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_PACK2=off MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_MARCH=lockstep MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 hires
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 stereo
done 2>&1 | grep -v Warning | tee gpurun_out/g_times.log
MD2_LIB_PATH=$L/libmd2loss_a2c4.so timeout 120 python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/g_times.log
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
export MD2_LIB_PATH=$L/libmd2loss_a2c4.so
$CMD > gpurun_out/g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_g_march $CMD > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
