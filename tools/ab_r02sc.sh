#!/bin/bash
# Round 2 (second session) third A/B pass: twiddle columns in tensor memory (main) against the same library without (notmem) and the
# round-2 library (prev), hybrid sizes; GPU parity suite first.
tag=${TAG:-r02sc}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
SIZES="${SIZES:-2048 4096 8192}" tools/ab_all.sh ${@:-prev notmem main} 2>&1 | tee gpurun_out/${tag}_ab.txt
