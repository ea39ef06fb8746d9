#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for v in a2c4 a1c5; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_PACK2=off MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_MARCH=lockstep MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 mono
  MD2_LIB_PATH=$L/libmd2loss_$v.so timeout 120 python scripts/time_loss.py 0 30 stereo
done 2>&1 | grep -v Warning | tee gpurun_out/h_times.log
export MD2_LIB_PATH=$L/libmd2loss_a2c4.so
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/h_pytest.log
tail -3 gpurun_out/h_pytest.log
