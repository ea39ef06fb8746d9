#!/bin/bash
cd "$(dirname "$0")/.."
L=monodepth2_b200/lib
for v in libmd2loss.so libmd2loss_v1.so libmd2loss_v3.so libmd2loss_v4.so; do
  for wl in mono stereo; do MD2_LIB_PATH=$L/$v timeout 120 python scripts/time_loss.py 0 30 $wl; done
done 2>&1 | grep -v Warning | tee gpurun_out/l_times.log
timeout 600 python -m pytest tests/test_gpu_mixin.py tests/test_gpu_parity.py -q -x > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log | cut -c1-200
timeout 900 python bench.py --no-cpu > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/l_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/l_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','value_cabi_predrawn_noise','e2e']}); print(d['train']['ms_per_step'])
PY
