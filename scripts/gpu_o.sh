#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
for v in libmd2loss.so libmd2loss_pc.so; do
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono
done
for v in libmd2loss_c6.so libmd2loss_c6pc.so; do
  for rows in 0 48 64; do MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py $rows 30 mono; done
done
done 2>&1 | grep -v Warning | tee gpurun_out/o_times.log
for v in libmd2loss.so libmd2loss_pc.so; do
  for wl in hires; do MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 $wl; done
done 2>&1 | grep -v Warning | tee -a gpurun_out/o_times.log
