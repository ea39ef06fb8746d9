#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for v in a1c5 a1c4; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_PACK2=off MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 stereo
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 hires
done 2>&1 | grep -v Warning | tee gpurun_out/i_times.log
MD2_LIB_PATH=$L/libmd2loss_a1c5.so timeout 120 python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/i_times.log
export MD2_LIB_PATH=$L/libmd2loss_a1c5.so
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/i_pytest.log
tail -15 gpurun_out/i_pytest.log
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
$CMD > gpurun_out/i_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_i_march $CMD > gpurun_out/i_ncu.log 2>&1
tail -2 gpurun_out/i_ncu.log
