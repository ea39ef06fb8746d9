#!/bin/bash
# round 2, call A: parity of the role kernel + timing of the variants
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
L=monodepth2_b200/lib
for wl in mono stereo hires; do
  MD2_MARCH=warp python scripts/time_loss.py 0 30 $wl
  for v in libmd2loss.so libmd2loss_r5.so libmd2loss_r6.so; do
    MD2_LIB_PATH=$L/$v python scripts/time_loss.py 0 30 $wl
  done
done 2>&1 | grep -v Warning | tee gpurun_out/a_times.log
for r in 32 48 64 96 192; do MD2_LIB_PATH=$L/libmd2loss_r5.so python scripts/time_loss.py $r 30 mono; done 2>&1 | grep -v Warning | tee -a gpurun_out/a_times.log
MD2_LIB_PATH=$L/libmd2loss_r5.so python scripts/time_loss.py 0 30 mono structured 2>&1 | grep -v Warning | tee -a gpurun_out/a_times.log
MD2_LIB_PATH=$L/libmd2loss_r5.so python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/a_times.log
MD2_MARCH=warp python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/a_times.log
