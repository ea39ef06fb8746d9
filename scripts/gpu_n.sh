#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/n_pytest.log
tail -8 gpurun_out/n_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n_bench.err; cat gpurun_out/n_bench.json | cut -c1-4000
