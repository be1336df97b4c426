#!/usr/bin/env python
"""Decisions per second for R co-located streaming radios: one launch per round (crn_create_many + crn_submit_many)
against today's one launch per radio (R crn_create handles, crn_submit each).  Every round each radio's K frames are
copied into its pinned ring slot, committed, and all R decisions are waited for.

  python tools/many_radios.py [--radios 256] [--mode ref|welch] [--nfft 1024] [--navg 64] [--rounds 50]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--radios", type=int, default=256)
    ap.add_argument("--mode", default="ref")
    ap.add_argument("--nfft", type=int, default=1024)
    ap.add_argument("--navg", type=int, default=64)
    ap.add_argument("--rounds", type=int, default=50)
    a = ap.parse_args()
    import numpy as np
    import crn_b200 as crn
    cfg = crn.config_reference() if a.mode == "ref" else crn.config_welch(a.nfft, a.navg)
    K, L, R = cfg.navg, cfg.frame_len, a.radios
    rng = np.random.default_rng(1)
    frames = (rng.standard_normal((R, K * L, 2)) * 0.1).astype(np.float32)   # one decision's frames per radio
    nbytes = K * L * 8

    def run(sensors, many):
        lat = []
        fill = 0.0
        t0 = time.perf_counter()
        for _ in range(a.rounds):
            f0 = time.perf_counter()
            for r, s in enumerate(sensors):
                C.memmove(s.ring_slot(), frames[r].ctypes.data, nbytes)
            f1 = time.perf_counter()
            fill += f1 - f0
            if many:
                crn.Sensor.submit_many(sensors, K)
            else:
                for s in sensors:
                    crn._check(crn.lib.crn_submit(s._h, K), "crn_submit")
            res = [s.wait() for s in sensors]
            lat.append(time.perf_counter() - f1)
        total = time.perf_counter() - t0
        return {"decisions_per_s": R * a.rounds / total, "decisions_per_s_excluding_fill": R * a.rounds / (total - fill),
                "submit_to_last_result_ms_median": 1e3 * sorted(lat)[len(lat) // 2], "fill_ms_per_round": 1e3 * fill / a.rounds,
                "launches": sensors[0].launches}, res

    pooled = crn.Sensor.create_many(cfg, R, device=0)
    run(pooled[:], True)
    b0 = pooled[0].launches
    out_many, res_many = run(pooled, True)
    out_many["launches"] -= b0
    for s in pooled:
        s.close()
    single = [crn.Sensor(cfg, device=0) for _ in range(R)]
    run(single, False)
    out_single, res_single = run(single, False)
    out_single["launches"] = sum(s.launches for s in single) // 2
    same = all(np.allclose(np.array(x.feat[:cfg.nbands]), np.array(y.feat[:cfg.nbands]), rtol=2e-6, atol=0) and x.decision == y.decision
               for x, y in zip(res_many, res_single))
    for s in single:
        s.close()
    print(json.dumps({"radios": R, "mode": a.mode, "nfft": cfg.nfft, "navg": K, "frame_len": L, "rounds": a.rounds,
                      "bytes_per_decision": nbytes, "one_launch_per_round": out_many, "one_launch_per_radio": out_single,
                      "speedup_excluding_fill": out_many["decisions_per_s_excluding_fill"] / out_single["decisions_per_s_excluding_fill"],
                      "same_results": bool(same)}))


if __name__ == "__main__":
    main()
