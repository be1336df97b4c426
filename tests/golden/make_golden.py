"""Generate the golden fixtures in this directory from oracle O1 = the reference's UNMODIFIED engine
object (/root/reference/cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp compiled into
oracle/_ref/libcrn_ref.so by oracle/Makefile).  Run in the build container, where /root/reference
exists:   python tests/golden/make_golden.py

The reference ships no golden vectors of its own: these are outputs of the reference's own
code run here on committed inputs, which is what pins the C port (oracle/crn_oracle.c) and, through it,
the CUDA path.  Inputs come from the seeded synthetic PU generator (oracle.synth) plus two analytic
cases (all zeros, one tone)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import crn_b200 as crn  # noqa: E402  (config structs only; no GPU call)
import oracle  # noqa: E402


def case(name, iq, L):
    feat, ann, dec, txf, bins = oracle.sense_ref(iq, L=L, want_bins=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), iq=iq.astype(np.complex64), L=np.int32(L),
                        feat=feat, ann=ann, decision=dec, tx_freq=txf, avg_bins=bins)
    print(name, "L=%d decisions=%s tx=%s" % (L, dec.tolist(), (txf / 1e6).tolist()))


def main():
    assert oracle.ref() is not None, "build oracle/_ref first (make -C oracle)"
    # 1. Markov-PU OFDM + AWGN, full frames (L = 512), 8 decisions, one hop per decision
    gs = 512 * 10
    sc = crn.synth_config(gs, dwell_groups=1, snr_db=10.0, seed=12)
    iq, states = oracle.synth(sc, 8 * gs)
    print("pu states", states.tolist())
    case("ref_markov_L512", iq, 512)
    # 2. ragged frames: L = 363 samples per packet (GigE USRP at MTU 1500), zero padded to 512 by the
    #    engine's memcpy into a zeroed buffer (CE_Predictive_Node.cpp:37,149)
    gs = 363 * 10
    sc = crn.synth_config(gs, dwell_groups=1, snr_db=5.0, seed=7)
    iq, _ = oracle.synth(sc, 6 * gs)
    case("ref_markov_L363", iq, 363)
    # 3. low SNR / noise only region: -5 dB
    gs = 512 * 10
    sc = crn.synth_config(gs, dwell_groups=2, snr_db=-5.0, seed=3)
    iq, _ = oracle.synth(sc, 6 * gs)
    case("ref_markov_snr-5", iq, 512)
    # 4. analytic: all zeros (ANN known answer) and one unit tone at bin 70 (inside CH2)
    case("ref_zeros", np.zeros(512 * 10, np.complex64), 512)
    n = np.arange(512 * 10)
    case("ref_tone70", np.exp(2j * np.pi * 70 * n / 512).astype(np.complex64), 512)


if __name__ == "__main__":
    main()
