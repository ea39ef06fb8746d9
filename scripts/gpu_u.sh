#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
for v in _m84 _m52 _m62; do MD2_LIB_PATH=$L/libmd2loss$v.so timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn; done
MD2_MARCH=lockstep MD2_LIB_PATH=$L/libmd2loss_m84.so timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn
done | tee gpurun_out/u_times.log
