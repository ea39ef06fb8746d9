#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for v in a2c4r a1c5r; do
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 mono
  MD2_PACK2=off MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 mono
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 hires
  MD2_LIB_PATH=$L/libmd2loss_$v.so python scripts/time_loss.py 0 30 stereo
done 2>&1 | grep -v Warning | tee gpurun_out/e_times.log
for r in 48 64; do MD2_LIB_PATH=$L/libmd2loss_a2c4r.so python scripts/time_loss.py $r 30 mono; done 2>&1 | grep -v Warning | tee -a gpurun_out/e_times.log
MD2_LIB_PATH=$L/libmd2loss_a2c4r.so python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/e_times.log
CMD="python scripts/time_loss.py 0 3 mono"
export MD2_LIB_PATH=$L/libmd2loss_a2c4r.so
$CMD > gpurun_out/e_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_e_march $CMD > gpurun_out/e_ncu.log 2>&1
tail -2 gpurun_out/e_ncu.log
