#!/bin/bash
# round 2, call B: packed role kernel: parity, timing of variants, ncu of the march kernel
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/b_pytest.log
tail -3 gpurun_out/b_pytest.log
L=monodepth2_b200/lib
for wl in mono; do
  for v in libmd2loss.so libmd2loss_r5.so; do
    MD2_LIB_PATH=$L/$v python scripts/time_loss.py 0 30 $wl
    MD2_PACK2=off MD2_LIB_PATH=$L/$v python scripts/time_loss.py 0 30 $wl
  done
done 2>&1 | grep -v Warning | tee gpurun_out/b_times.log
for r in 48 64 96; do MD2_LIB_PATH=$L/libmd2loss.so python scripts/time_loss.py $r 30 mono; done 2>&1 | grep -v Warning | tee -a gpurun_out/b_times.log
MD2_LIB_PATH=$L/libmd2loss.so python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/b_times.log
MD2_LIB_PATH=$L/libmd2loss.so python scripts/time_loss.py 0 30 hires 2>&1 | grep -v Warning | tee -a gpurun_out/b_times.log
MD2_LIB_PATH=$L/libmd2loss_r5.so python scripts/time_loss.py 0 30 hires 2>&1 | grep -v Warning | tee -a gpurun_out/b_times.log
CMD="python scripts/time_loss.py 0 3 mono"
$CMD > gpurun_out/b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_b_march $CMD > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
