#!/bin/bash
# round 2, call 15 (8 GPUs): in-kernel timeline of the sweeps at 8 and 1 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2q_trace8.rank* gpurun_out/r2q_trace1.rank*
MF_SWEEP_TRACE=$PWD/gpurun_out/r2q_trace8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --legs none --no-bitwise-check --no-launch-timing > gpurun_out/r2q_8.json 2> gpurun_out/r2q_8.err; echo "exit $?"
for r in 0 3 7; do echo "rank $r"; python scripts/trace_summary.py gpurun_out/r2q_trace8.rank$r 480; done
MF_SWEEP_TRACE=$PWD/gpurun_out/r2q_trace1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --legs none --no-launch-timing > gpurun_out/r2q_1.json 2> gpurun_out/r2q_1.err; echo "exit $?"
python scripts/trace_summary.py gpurun_out/r2q_trace1.rank0 480
gzip -f gpurun_out/r2q_trace8.rank* gpurun_out/r2q_trace1.rank*
