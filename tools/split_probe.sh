for ng in 512 1024 2048 4096; do
  samples=$((ng*131072))
  for sp in 1 2 4 8; do
    echo "ng=$ng split=$sp :: $(CRN_FORCE_SPLIT=$sp python tools/kbench.py --nfft 2048 --mode welch --samples $samples --steps 50 --reps 3 2>&1 | tail -1 | cut -c40-110)"
  done
done
