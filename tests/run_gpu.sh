#!/bin/bash
# GPU-box helper: every stage under its own timeout, logs under gpurun_out/ (which gpurun brings back).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
