#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for c in -1 48 55 65 80 100; do echo "carveout $c: $(MD2_CARVEOUT=$c MD2_LIB_PATH=$L/libmd2loss_kn.so timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn | awk '{print $8,$9,$10,$11}')"; done | tee gpurun_out/u_times.log
