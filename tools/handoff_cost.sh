#!/bin/bash
# Host-side cost of the rx -> CE -> libcrnsense handoff per packet, received in place vs copied twice as upstream does
# (tools/handoff_cost.sh, GPU box): crn_replay loops a capture for 200000 packets in lock-step.
python - <<'PY'
import numpy as np
np.load("tests/golden/ref_markov_L512.npz")["iq"].astype(np.complex64).tofile("/tmp/cap512.c64")
(np.random.default_rng(0).standard_normal((64 * 8192, 2)).astype(np.float32) * 0.1).tofile("/tmp/cap8192.c64")
PY
R=cognitive-radio-network_b200/host/crn_replay
for mode in direct copy; do
  for cfg in "512 -d 0 -q" "8192 -n 8192 -k 64 -w 1 -p 1 -d 0 -q"; do
    set -- $cfg; L=$1; shift
    if [ $mode = copy ]; then export CRN_ENGINE_COPY=1; else unset CRN_ENGINE_COPY; fi
    echo "$mode L=$L :: $($R --scenario tests/golden/predictive_model_su.cfg --node 2 --iq /tmp/cap$L.c64 --packet-len $L --repeat-packets 200000 --ce-args "$*" | tail -1)"
  done
done
