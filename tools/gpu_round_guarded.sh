#!/bin/bash
# Parity suite first under a short timeout; the full profile pass only if it is green (protects the GPU budget).
tag=${1:-r02sm}
mkdir -p gpurun_out
if ! timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.txt 2>&1; then
  tail -15 gpurun_out/${tag}_pytest.txt; echo "GPU TESTS NOT GREEN - stopping"; exit 1
fi
tail -2 gpurun_out/${tag}_pytest.txt
SIZES="1024" STEPS=10 tools/ab_all.sh prev main 2>&1 | tee gpurun_out/${tag}_ab.txt
timeout 900 tools/gpu_round.sh $tag
