#!/bin/bash
# One gpurun call: GPU tests, smoke, a short bench, then (only if the bench passed) the ncu launch list.
# Usage: scripts/gpu_check.sh [pytest-args...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 "$@" > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
