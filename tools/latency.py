#!/usr/bin/env python
"""Decision latency of the streaming path (pinned ring -> H2D -> fused kernel -> D2H, crn_submit .. crn_wait):
what one CE_Predictive_Node instance sees per decision when frames arrive one at a time
(src/extensible_cognitive_radio.cpp:1310-1324 handoff).   python tools/latency.py [--reps 200]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    import crn_b200 as crn
    rows = []
    rng = np.random.default_rng(1)
    for name, cfg in (("reference-exact N=512 K=10", crn.config_reference()),
                      ("welch N=1024 K=64", crn.config_welch(1024, 64)),
                      ("wideband N=8192 K=64", crn.config_wideband(8192, 64, 64))):
        L, K = cfg.frame_len, cfg.navg
        frames = (rng.standard_normal((K, L)) + 1j * rng.standard_normal((K, L))).astype(np.complex64) * 0.1
        with crn.Sensor(cfg, device=0) as s:
            for _ in range(5):                       # warm-up decisions
                for k in range(K):
                    s.push_frame(frames[k])
                s.wait()
            last, whole = [], []
            for _ in range(a.reps):
                t0 = time.perf_counter()
                for k in range(K - 1):
                    s.push_frame(frames[k])
                t1 = time.perf_counter()
                s.push_frame(frames[K - 1])          # K-th frame: ships the slot and launches the kernel
                s.wait()
                t2 = time.perf_counter()
                last.append(t2 - t1)
                whole.append(t2 - t0)
        last.sort()
        whole.sort()
        rows.append({"config": name, "bytes_per_decision": 8 * K * L,
                     "kth_frame_to_decision_us": {"median": 1e6 * last[len(last) // 2], "p99": 1e6 * last[int(0.99 * len(last))]},
                     "decisions_per_s_one_stream": 1.0 / whole[len(whole) // 2]})
        print("%-28s %8d B/decision   K-th frame -> decision: median %7.1f us  p99 %7.1f us   %8.0f decisions/s (one stream, python producer)"
              % (name, 8 * K * L, 1e6 * last[len(last) // 2], 1e6 * last[int(0.99 * len(last))], 1.0 / whole[len(whole) // 2]))
    print(json.dumps({"streaming_latency": rows}))


if __name__ == "__main__":
    main()
