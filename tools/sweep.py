#!/usr/bin/env python
"""FFT-size sweep (BASELINE configs[2..4] shapes and the configs[4] sweep): device-resident throughput of
the fused sensing kernel per N, as Gsamples/s and fraction of the measured HBM roofline.

  python tools/sweep.py [--samples 2.5e8] [--steps 10] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=float, default=2.5e8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch
    import crn_b200 as crn
    torch.cuda.set_device(0)
    peak = 6543.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    stream = torch.cuda.current_stream().cuda_stream
    nsamp = int(args.samples)
    d_iq = torch.empty(nsamp, 2, dtype=torch.float32, device="cuda")
    sc = crn.synth_config(65536, dwell_groups=64, snr_db=10.0, seed=12)
    crn.synth_generate(sc, d_iq, 0, nsamp, None, 0, stream)
    rows = []
    cases = []
    for n in (256, 512, 1024, 2048, 4096, 8192):
        cases.append(("welch K=64", crn.config_welch(n, 64) if n >= 512 else None))
        cases.append(("wideband 64ch K=64", crn.config_wideband(n, 64, 64 if n >= 512 else 16)))
    cases.append(("reference-exact", crn.config_reference()))
    for name, cfg in cases:
        if cfg is None:
            continue
        gs = cfg.group_samples
        ng = nsamp // gs
        with crn.Sensor(cfg, device=0) as s:
            info = s.kernel_info()
            d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
            d_ann = torch.empty(ng, 3, dtype=torch.float64, device="cuda")
            d_dec = torch.empty(ng, dtype=torch.int32, device="cuda")
            d_mask = torch.empty(ng, dtype=torch.int64, device="cuda")
            for _ in range(3):
                s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, d_mask, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, d_mask, stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        gsps = ng * gs / ms / 1e6
        row = {"nfft": cfg.nfft, "mode": name, "navg": cfg.navg, "groups": ng, "ms": ms, "gsamples_s": gsps,
               "hbm_gbs": gsps * 8, "frac_of_measured_hbm": gsps * 8 / peak, "kernel": info["name"],
               "regs": info["regs_per_thread"], "ctas_per_sm": info["ctas_per_sm"], "smem": info["smem_bytes"],
               "threads_per_cta": info["threads_per_cta"]}
        rows.append(row)
        print("N=%5d %-20s %8.1f GS/s  %6.1f%% of measured HBM  (%s, %d regs, %d CTA/SM x %d thr, %d B smem)" %
              (cfg.nfft, name, gsps, 100 * row["frac_of_measured_hbm"], info["name"], info["regs_per_thread"],
               info["ctas_per_sm"], info["threads_per_cta"], info["smem_bytes"]), flush=True)
    if args.json:
        json.dump({"samples": nsamp, "steps": args.steps, "hbm_peak_gbs": peak, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
