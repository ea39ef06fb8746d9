#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
for v in libmd2loss.so libmd2loss_vJ.so libmd2loss_vK.so libmd2loss_vM.so; do
  for wl in mono; do MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 $wl; done
done; done 2>&1 | grep -v Warning | tee gpurun_out/m_times.log
for v in libmd2loss.so libmd2loss_vJ.so libmd2loss_vK.so; do
  for wl in stereo hires; do MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 $wl; done
done 2>&1 | grep -v Warning | tee -a gpurun_out/m_times.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv | tee -a gpurun_out/m_times.log
