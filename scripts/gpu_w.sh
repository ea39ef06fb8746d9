#!/bin/bash
# A/B on one box: the library before the scale-level change (libmd2loss_prev.so) against the shipped one; then the GPU tests
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for rep in 1 2; do
  for lib in libmd2loss_prev.so libmd2loss.so; do
    for wl in mono stereo hires; do
      MD2_LIB_PATH=$L/$lib timeout 120 python scripts/time_loss.py 0 40 $wl 2>&1 | grep -v Warn
    done
  done
done | tee gpurun_out/w_ab2.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02w_pytest.log
tail -4 gpurun_out/r02w_pytest.log | cut -c1-300
