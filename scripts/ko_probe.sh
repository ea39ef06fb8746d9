#!/bin/bash
# development aid: A/B timing of kernel variants
L=monodepth2_b200/lib
for lib in libmd2_bs3.so libmd2_p3.so; do
  for wl in mono; do
    MD2_LIB_PATH=$L/$lib python scripts/time_loss.py 0 30 $wl iid grad
  done
done
