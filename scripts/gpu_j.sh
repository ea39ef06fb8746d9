#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/j_pytest.log
tail -6 gpurun_out/j_pytest.log | cut -c1-200
for wl in mono stereo hires; do timeout 120 python scripts/time_loss.py 0 30 $wl; done 2>&1 | grep -v Warning | tee gpurun_out/j_times.log
timeout 120 python scripts/time_loss.py 0 30 mono iid nograd 2>&1 | grep -v Warning | tee -a gpurun_out/j_times.log
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/j_bench.err; cat gpurun_out/j_bench.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/j_bench_ref.json | cut -c1-800
