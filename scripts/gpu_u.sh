#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/u_pytest.log
for wl in mono_640x192_b12 mono_640x192_b12_avg_reprojection; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value %.0f'%d['value'], 'march %.4f'%d['roofline']['kernel_ms'], 'e2e %.0f'%d['e2e']['value'])"
done | tee gpurun_out/u_times.log
