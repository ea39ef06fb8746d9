#!/bin/bash
# development aid: A/B timing of kernel variants
L=monodepth2_b200/lib
for lib in libmd2_s3.so libmd2_s2.so; do
  MD2_PACK2=all MD2_LIB_PATH=$L/$lib python scripts/time_loss.py 0 30 mono iid grad
done
