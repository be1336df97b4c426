#!/bin/bash
# Round check on the GPU box: smoke, GPU parity tests, run-to-run stress, compute-sanitizer, default bench.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 900 python tests/stress_gpu.py 2>&1 | tail -20
timeout 1500 bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; tail -60 gpurun_out/sanitize.log
python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; cat gpurun_out/bench_check.json
