#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do for v in "" _a2; do MD2_LIB_PATH=$L/libmd2loss$v.so timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn; done; done | tee gpurun_out/u_times.log
for v in "" _a2; do for wl in hires; do MD2_LIB_PATH=$L/libmd2loss$v.so timeout 120 python scripts/time_loss.py 0 30 $wl 2>&1 | grep -v Warn; done; done | tee -a gpurun_out/u_times.log
MD2_LIB_PATH=$L/libmd2loss_a2.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu 2>&1 | tail -2
