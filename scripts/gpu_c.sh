#!/bin/bash
# round 2, call C: role kernel with shared role A + B prefetch
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c_pytest.log
tail -3 gpurun_out/c_pytest.log
L=monodepth2_b200/lib
for v in a2c4 a2c3 a1c5 a3c3; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 mono
  MD2_PACK2=off MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 mono
done 2>&1 | grep -v Warning | tee gpurun_out/c_times.log
for v in a2c4 a2c3 a3c3; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 stereo
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 hires
done 2>&1 | grep -v Warning | tee -a gpurun_out/c_times.log
for r in 48 64 96; do MD2_LIB_PATH=$L/libmd2loss_a2c4.so python scripts/time_loss.py $r 30 mono; done 2>&1 | grep -v Warning | tee -a gpurun_out/c_times.log
MD2_LIB_PATH=$L/libmd2loss_a2c4.so python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/c_times.log
CMD="python scripts/time_loss.py 0 3 mono"
$CMD > gpurun_out/c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_c_march $CMD > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
