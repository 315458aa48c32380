#!/usr/bin/env bash
# Multi-GPU session on ONE box with N GPUs (charged N x): sharded parity tests over world 2/4/8, the
# bench line at 1/2/4/N GPUs (the driver's torchrun launch), one kernel timeline at N GPUs.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_session_multi.sh 8'
set -u
N=${1:-8}
PORT=${PORT:-29517}
mkdir -p gpurun_out
echo "== sharded parity (world 2/4/8 + ragged shards; eager, graphs, two-pass, NCCL fallback)" | tee gpurun_out/session_multi.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4 | tee -a gpurun_out/session_multi.log
for G in 1 2 4 $N; do
  [ "$G" -gt "$N" ] && continue
  echo "== bench at $G GPUs" | tee -a gpurun_out/session_multi.log
  if [ "$G" -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > "gpurun_out/bench_${G}gpu.json" 2> "gpurun_out/bench_${G}gpu.err"
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$G" --master-addr 127.0.0.1 \
        --master-port "$PORT" bench.py --gpus "$G" --steps 20 --warmup 5 \
        > "gpurun_out/bench_${G}gpu.json" 2> "gpurun_out/bench_${G}gpu.err"
  fi
  echo "rc=$?" | tee -a gpurun_out/session_multi.log
  python - "$G" <<'PY' | tee -a gpurun_out/session_multi.log
import json, sys
try:
    l = json.load(open(f"gpurun_out/bench_{sys.argv[1]}gpu.json"))
    print({k: l[k] for k in ("n_gpus", "ms_per_step", "value", "pct_of_bf16_peak")}, "parity", l["parity"]["ok"],
          {k: f"{l['parity'][k]:.1e}" for k in ("loss_rel_err", "dI_rel_err", "dT_rel_err", "dlogit_scale_rel_err")},
          "e2e", round(l["e2e"]["ms_per_step"], 3), "head", round(l["full_head"]["ms_per_step"], 3), "weak", round(l["weak_scaling_point"]["ms_per_step"], 3))
except Exception as e:
    print("no bench line:", e)
PY
  PORT=$((PORT + 1))
done
echo "== kernel timeline of one step at $N GPUs" | tee -a gpurun_out/session_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((PORT + 7)) tools/trace_step_dist.py > "gpurun_out/trace_${N}gpu.txt" 2> "gpurun_out/trace_${N}gpu.err"
tail -3 "gpurun_out/trace_${N}gpu.txt" | tee -a gpurun_out/session_multi.log
