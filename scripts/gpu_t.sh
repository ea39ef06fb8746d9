#!/bin/bash
# usage: gpu_t.sh <tag> <variant> [<variant> ...]   times libmd2loss.so and each libmd2loss_<variant>.so (march kernel alone),
# runs the parity tests on the LAST variant and captures one ncu profile of its marching kernel
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
tag=$1; shift
last=""
for rep in 1 2; do
for v in "" "$@"; do
  lib=libmd2loss${v:+_$v}.so
  MD2_LIB_PATH=$L/$lib timeout 120 python scripts/time_loss.py 0 30 mono
done; done 2>&1 | grep -v Warning | tee gpurun_out/${tag}_times.log
for v in "$@"; do last=$v; done
MD2_LIB_PATH=$L/libmd2loss_$last.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.log
CMD="timeout 200 python scripts/time_loss.py 0 3 mono"
MD2_LIB_PATH=$L/libmd2loss_$last.so ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_${tag}_$last $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_ncu.log
