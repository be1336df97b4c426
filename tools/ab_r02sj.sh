#!/bin/bash
# Round 2 (second session): per-band segment ranges in the epilogue (small sizes, 64 sub-channels), TMEM rows with / without a row in
# flight at 8192; then compute-sanitizer over every kernel family.
tag=${TAG:-r02sj}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
{
SIZES="512 1024" tools/ab_all.sh prev main
SIZES="2048 4096 8192" tools/ab_all.sh nopipe main
} 2>&1 | tee gpurun_out/${tag}_ab.txt
tools/sanitize.sh > gpurun_out/${tag}_sanitize.txt 2>&1
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/${tag}_sanitize.txt; grep "SUMMARY" gpurun_out/${tag}_sanitize.txt | sort | uniq -c
