#!/bin/bash
cd "$(dirname "$0")/.."
for wl in mono stereo hires; do timeout 120 python scripts/time_loss.py 0 30 $wl 2>&1 | grep -v Warn; done | tee gpurun_out/u_times.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/u_pytest.log
timeout 600 python bench.py --no-cpu --no-train > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/u_bench.err
python - <<PY
import json
for f in ['gpurun_out/u_bench.json']:
    d=json.load(open(f)); r=d['roofline']
    print(f, 'value %.0f ms %.4f cabi %.0f e2e %.0f march %.4f frac %.3f'%(d['value'], d['ms_per_step'], d.get('value_cabi_predrawn_noise',0), d['e2e']['value'], r['kernel_ms'], r['frac']))
PY
