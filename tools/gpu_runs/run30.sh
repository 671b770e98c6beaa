#!/bin/bash
# 1 GPU, final build: the whole GPU suite + smoke
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
