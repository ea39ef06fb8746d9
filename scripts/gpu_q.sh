#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
for v in libmd2loss.so "$@"; do
  MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 mono
done; done 2>&1 | grep -v Warning | tee gpurun_out/q_times.log
for v in "$@"; do
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
MD2_LIB_PATH=$L/$v ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_q_${v%.so} $CMD > gpurun_out/q_ncu.log 2>&1
tail -1 gpurun_out/q_ncu.log
done
