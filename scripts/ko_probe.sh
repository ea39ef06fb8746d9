#!/bin/bash
# development aid: A/B timing of kernel variants
L=monodepth2_b200/lib
for lib in libmd2loss.so libmd2_bs.so libmd2_pinbs.so libmd2_pin.so; do
  MD2_LIB_PATH=$L/$lib python scripts/time_loss.py 0 30 mono iid grad
done
MD2_LIB_PATH=$L/libmd2_bs.so python scripts/time_loss.py 0 30 hires iid grad
MD2_LIB_PATH=$L/libmd2_bs.so python scripts/time_loss.py 0 30 mono structured grad
