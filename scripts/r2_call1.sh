#!/bin/bash
# round 2, call 1: GPU tests (incl. the config-scale parity tests), headline bench with the new legs, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; df -h /tmp >> gpurun_out/r2a_host.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=20 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2a_pytest.log
MF_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; cat gpurun_out/r2a_bench.json | cut -c1-6000
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref exit $?"; cat gpurun_out/r2a_ref.json | cut -c1-3000
