#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for v in "" _f24 _f32 _f8; do
python bench.py --steps 2 --warmup 1 --no-cpu --no-graph --no-train > /dev/null 2>&1
MD2_LIB_PATH=$L/libmd2loss$v.so ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_u$v.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-graph --no-train > /dev/null 2>&1
echo "final$v: $(grep 'md2_final' gpurun_out/launches_u$v.csv | tail -2 | awk -F'\",\"' '{print $(NF)}' | tr '\n' ' ')"
MD2_LIB_PATH=$L/libmd2loss$v.so timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn | awk '{print $8,$9}'
done
MD2_LIB_PATH=$L/libmd2loss_f24.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -2
