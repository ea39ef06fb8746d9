#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_color_jitter.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/u_pytest_cj.log
