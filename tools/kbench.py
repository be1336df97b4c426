#!/usr/bin/env python
"""Quick device-resident timing of one configuration (kernel development loop).
   python tools/kbench.py [--nfft 1024] [--navg 64] [--mode welch|wide|ref] [--samples 1e9] [--steps 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=1024)
    ap.add_argument("--navg", type=int, default=64)
    ap.add_argument("--mode", default="welch")
    ap.add_argument("--samples", type=float, default=1e9)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--window", default=None)
    a = ap.parse_args()
    import torch
    import crn_b200 as crn
    if a.mode == "welch":
        cfg = crn.config_welch(a.nfft, a.navg)
    elif a.mode == "wide":
        cfg = crn.config_wideband(a.nfft, a.navg, 64 if a.nfft >= 512 else 16)
    else:
        cfg = crn.config_reference()
    if a.window == 'rect':
        cfg.window = crn.WINDOW_RECT
    gs = cfg.group_samples
    ng = int(a.samples) // gs
    stream = torch.cuda.current_stream().cuda_stream
    d_iq = torch.empty(ng * gs, 2, dtype=torch.float32, device="cuda")
    crn.synth_generate(crn.synth_config(gs, dwell_groups=64), d_iq, 0, ng * gs, None, 0, stream)
    d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
    d_ann = torch.empty(ng, 3, dtype=torch.float64, device="cuda")
    d_dec = torch.empty(ng, dtype=torch.int32, device="cuda")
    d_mask = torch.empty(ng, dtype=torch.int64, device="cuda")
    with crn.Sensor(cfg) as s:
        info = s.kernel_info()
        for _ in range(3):
            s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, d_mask, stream)
        best = 1e9
        for _ in range(a.reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, d_mask, stream)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / a.steps)
    g = ng * gs / best / 1e6
    print("%s N=%d K=%d groups=%d: %.3f ms  %.1f GS/s  %.1f GB/s  %.3f of 6543.4  (regs %d, %d CTA/SM, smem %d)  checksum %.6e" %
          (info["name"], cfg.nfft, cfg.navg, ng, best, g, g * 8, g * 8 / 6543.4, info["regs_per_thread"],
           info["ctas_per_sm"], info["smem_bytes"], d_feat.double().sum().item()))


if __name__ == "__main__":
    main()
