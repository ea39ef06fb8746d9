#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r_pytest.log
tail -8 gpurun_out/r_pytest.log | cut -c1-300
for wl in mono stereo hires; do timeout 120 python scripts/time_loss.py 0 30 $wl; done 2>&1 | grep -v Warning | tee gpurun_out/r_times.log
