#!/bin/bash
# Final GPU pass of round 2 (second session): everything tools/gpu_round.sh produces, then racecheck of the 2048 / 4096 kernels with and
# without the mbarrier hand-off (racecheck models bar.sync only: the hazards it lists must vanish in the variant that
# keeps the team barrier).
tag=${1:-r02sk}
tools/gpu_round.sh $tag
V=$PWD/cognitive-radio-network_b200/variants
{
for lib in "" "CRN_LIB=$V/libcrnsense_noef.so"; do
  for args in "--nfft 2048 --navg 64 --samples 3e6" "--nfft 4096 --mode wide --samples 2e6"; do
    echo "== racecheck ${lib:+team-barrier variant (CRN_EARLY_FREE_MAXC=0) }$args"
    env $lib timeout 600 compute-sanitizer --tool racecheck --print-limit 2 python tools/kbench.py $args --steps 1 --reps 1 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|sense_n" | cut -c1-200 | head -4
  done
done
} > gpurun_out/${tag}_racecheck_handoff.txt 2>&1
cat gpurun_out/${tag}_racecheck_handoff.txt
