#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
V=$1
for rep in 1 2; do
for v in libmd2loss.so libmd2loss_$V.so; do
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono
done; done 2>&1 | grep -v Warning | tee gpurun_out/s_times.log
MD2_LIB_PATH=$L/libmd2loss_$V.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "mono or avg" 2>&1 | tail -2
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
MD2_LIB_PATH=$L/libmd2loss_$V.so ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_s_$V $CMD > gpurun_out/s_ncu.log 2>&1
tail -1 gpurun_out/s_ncu.log
