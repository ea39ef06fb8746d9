#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for wl in mono stereo hires; do
  timeout 120 python scripts/time_loss.py 0 30 $wl
  MD2_LIB_PATH=$L/libmd2loss_notma.so timeout 120 python scripts/time_loss.py 0 30 $wl
done 2>&1 | grep -v Warning | tee gpurun_out/k_times.log
timeout 120 python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/k_times.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/k_pytest.log
tail -6 gpurun_out/k_pytest.log | cut -c1-200
