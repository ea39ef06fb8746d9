#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
