#!/bin/bash
# ThreadSanitizer pass over the host runtime (CPU only, needs the reference tree): the UHD-free radio's rx / CE
# worker pair with the reference's unmodified CPU CE_Predictive_Node plugged in, lock-step and free-run.
# Usage: tools/tsan_host.sh [reference cognitive_engines dir]     Expect no "WARNING: ThreadSanitizer" lines.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${1:-/root/reference/cognitive_engines}
OUT=$(mktemp -d)
gcc -O1 -g -c "$ROOT/oracle/liquid_fft_restated.c" -o "$OUT/fft.o"
make -s -C "$ROOT/cognitive-radio-network_b200/host" OUT="$OUT" ENGINE_DIRS="$REF" \
     EXTRA_INC="-I $ROOT/oracle/compat -fsanitize=thread" EXTRA_SRCS="$OUT/fft.o" 2>/dev/null
python - "$ROOT" "$OUT" <<'PY'
import sys, numpy as np
g = np.load(sys.argv[1] + "/tests/golden/ref_markov_L363.npz")
g["iq"].astype(np.complex64).tofile(sys.argv[2] + "/cap.c64")
PY
for mode in "" "--free-run"; do
  TSAN_OPTIONS="halt_on_error=0" "$OUT/crn_replay" --scenario "$ROOT/tests/golden/predictive_model_su.cfg" --node 2 \
      --iq "$OUT/cap.c64" --packet-len 363 --ce-args "" $mode 2>&1 | grep -E "WARNING: ThreadSanitizer|packets received" | sort | uniq -c
done
rm -rf "$OUT"
