#!/bin/bash
# end-of-round measurement pass: GPU tests, the bench line of both arms, the other BASELINE workloads, profiles
cd "$(dirname "$0")/.."
tag=${1:-r02f}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
for wl in mono+stereo_640x192_b12 mono_1024x320_b12 mono_640x192_b12_avg_reprojection mono_640x192_b12_disable_automasking; do
  timeout 600 python bench.py --workload $wl --no-cpu --no-train > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; echo "$wl rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_bench*.json')):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    r=d.get('roofline') or {}
    print(f.split('/')[-1], 'value %.0f'%d['value'], 'ms %.4f'%d['ms_per_step'], 'cabi %.0f'%d.get('value_cabi_predrawn_noise',0), 'e2e %.0f'%(d.get('e2e') or {}).get('value',0), 'march_ms %.4f frac %.3f'%(r.get('kernel_ms',0), r.get('frac',0)), 'train', (d.get('train') or {}).get('ms_per_step'))
PY
bash scripts/profile_round.sh $tag > gpurun_out/${tag}_profile.log 2>&1; tail -3 gpurun_out/${tag}_profile.log
