#!/usr/bin/env bash
# Multi-GPU call (charged N x): 2-GPU NCCL parity test, the bench line at N GPUs, one kernel timeline.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_session_multi.sh 8'
set -u
N=${1:-8}
PORT=${PORT:-29517}
mkdir -p gpurun_out
echo "== 2-GPU parity (eager, graphs, both reduce-scatter paths)" | tee gpurun_out/session_multi.log
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/session_multi.log
echo "== bench at $N GPUs" | tee -a gpurun_out/session_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port "$PORT" bench.py --gpus "$N" --steps 30 --warmup 5 \
    > "gpurun_out/bench_${N}gpu.json" 2> "gpurun_out/bench_${N}gpu.err"
tail -c 700 "gpurun_out/bench_${N}gpu.json" | tee -a gpurun_out/session_multi.log
echo "== kernel timeline of one step at $N GPUs" | tee -a gpurun_out/session_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((PORT + 1)) tools/trace_step_dist.py > "gpurun_out/trace_${N}gpu.txt" 2> "gpurun_out/trace_${N}gpu.err"
tail -5 "gpurun_out/trace_${N}gpu.txt" | tee -a gpurun_out/session_multi.log
