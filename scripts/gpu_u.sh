#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/u_pytest.log
for wl in mono stereo hires; do timeout 120 python scripts/time_loss.py 0 30 $wl 2>&1 | grep -v Warn; done | tee gpurun_out/u_times.log
