#!/bin/bash
# round-2 last session: GPU tests of the scale-subset build, smoke, the bench line
cd "$(dirname "$0")/.."
tag=${1:-r02v}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
r=d.get('roofline') or {}
print('value %.0f'%d['value'], 'ms %.4f'%d['ms_per_step'], 'cabi %.0f'%d.get('value_cabi_predrawn_noise',0), 'e2e %.0f'%(d.get('e2e') or {}).get('value',0), 'march_ms %.4f frac %.3f'%(r.get('kernel_ms',0), r.get('frac',0)))
PY
