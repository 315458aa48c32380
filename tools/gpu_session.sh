#!/usr/bin/env bash
# First GPU call of a session: sanity of the shipped path, then every staged kernel variant.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'
# Everything lands in gpurun_out/ (merged back by gpurun).  Each step has its own timeout so that a
# hung experimental kernel cannot eat the call.
set -u
mkdir -p gpurun_out
N=${N:-32768}
D=${D:-512}
echo "== shipped path: GPU tests" | tee gpurun_out/session.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/session.log
echo "== shipped path: bench" | tee -a gpurun_out/session.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_shipped.json 2> gpurun_out/bench_shipped.err
tail -c 600 gpurun_out/bench_shipped.json | tee -a gpurun_out/session.log
echo "== staged variants (csrc/next): parity, kernel times, wait profiles" | tee -a gpurun_out/session.log
timeout 1200 python tools/pipeline_experiments.py run "$N" "$D" > gpurun_out/pipeline_experiments.stdout 2>&1
grep -E "parity|fwd rows-only|rank shape|TIMEOUT|rc=" gpurun_out/pipeline_experiments.txt | tee -a gpurun_out/session.log
