#!/bin/bash
# ptxas -regUsageLevel variants of the whole library against the shipped one, same box
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
  for lib in libmd2loss.so libmd2loss_ru0.so libmd2loss_ru3.so libmd2loss_ru7.so libmd2loss_ru10.so; do
    [ -f $L/$lib ] || continue
    for wl in mono stereo hires; do
      MD2_LIB_PATH=$L/$lib timeout 120 python scripts/time_loss.py 0 40 $wl 2>&1 | grep -v Warn
    done
  done
done | tee gpurun_out/x_regusage.log
