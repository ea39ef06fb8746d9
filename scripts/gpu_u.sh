#!/bin/bash
cd "$(dirname "$0")/.."
timeout 120 python scripts/time_loss.py 0 30 mono 2>&1 | grep -v Warn | tee gpurun_out/u_times.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_variants.py tests/test_gpu_fuzz.py -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/u_pytest.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-graph --no-train > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_u.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-graph --no-train > /dev/null 2>&1
grep "depth_up" gpurun_out/launches_u.csv | tail -2 | awk -F'","' '{print $5, $(NF)}'
timeout 600 python bench.py --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms %.4f e2e %.0f march %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']))"
