#!/usr/bin/env python
"""Occupancy decisions against the true primary-user state, per SNR: the reference's weight literals
(CE_Predictive_Node.cpp:78-120) vs a predictor retrained on the GPU (crn_ann_train_device) from labelled feature
vectors of the same synthetic capture family ("Array of features + label", Data Generation/TODO.md).
Everything runs on the device: crn_synth -> fused sensing kernel (features) -> trainer -> fused kernel again with
the new weights (decisions).   python tools/retrain_demo.py [--nfft 1024] [--navg 64] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

SNRS = (-10.0, -5.0, 0.0, 5.0, 10.0, 20.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=1024)
    ap.add_argument("--navg", type=int, default=64)
    ap.add_argument("--train", type=int, default=400, help="labelled decisions per SNR used for training")
    ap.add_argument("--test", type=int, default=2000, help="held-out decisions per SNR")
    ap.add_argument("--epochs", type=int, default=20000)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import torch
    import crn_b200 as crn
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    cfg = crn.config_welch(a.nfft, a.navg)
    gs = cfg.group_samples
    ng = a.train + a.test
    feats, labels = {}, {}
    caps = {}
    with crn.Sensor(cfg, device=0) as s:
        for snr in SNRS:
            sc = crn.synth_config(gs, dwell_groups=2, snr_db=snr, seed=100 + int(snr), hop_mode=2)
            d_iq = torch.empty(ng * gs, 2, dtype=torch.float32, device="cuda")
            d_state = torch.empty(ng, dtype=torch.int32, device="cuda")
            crn.synth_generate(sc, d_iq, 0, ng * gs, d_state, 0, stream)
            d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
            d_dec = torch.empty(ng, dtype=torch.int32, device="cuda")
            s.sense_device(d_iq, ng, d_feat, None, d_dec, None, stream)
            torch.cuda.synchronize()
            caps[snr] = d_iq
            feats[snr], labels[snr] = d_feat, (d_state + 1).to(torch.int32)
            caps[snr + 0.5] = d_dec  # decisions with the reference's weights
    # one predictor for all SNRs: train on the first `train` decisions of every capture
    tr_feat = torch.cat([feats[snr][: a.train] for snr in SNRS]).contiguous()
    tr_lab = torch.cat([labels[snr][: a.train] for snr in SNRS]).contiguous()
    scale = [1.0 / float(tr_feat[:, i].max()) for i in range(4)]
    tc = crn.ann_train_config(max_epochs=a.epochs, check_every=500, eta=0.5, alpha=0.9, input_scale=scale,
                              init_range=0.5, seed=12, target_error=0.002 * tr_feat.shape[0])
    w, err, epochs = crn.ann_train(tc, tr_feat, cfg.nbands, tr_lab, tr_feat.shape[0], stream=stream)
    cfg2 = w.into_config(cfg.copy())
    rows = []
    with crn.Sensor(cfg2, device=0) as s2:
        for snr in SNRS:
            d_dec2 = torch.empty(ng, dtype=torch.int32, device="cuda")
            d_feat2 = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
            s2.sense_device(caps[snr], ng, d_feat2, None, d_dec2, None, stream)
            torch.cuda.synchronize()
            truth = labels[snr][a.train:]
            acc_ref = float((caps[snr + 0.5][a.train:] == truth).float().mean())
            acc_new = float((d_dec2[a.train:] == truth).float().mean())
            rows.append({"snr_db": snr, "held_out_decisions": a.test, "accuracy_reference_weights": acc_ref,
                         "accuracy_retrained": acc_new})
            print("SNR %+5.0f dB   reference weights %.3f   retrained on the GPU %.3f" % (snr, acc_ref, acc_new))
    print("trainer: %d examples, %d epochs, final E = %.4f" % (tr_feat.shape[0], epochs, err))
    if a.json:
        json.dump({"nfft": a.nfft, "navg": a.navg, "train_per_snr": a.train, "epochs": epochs, "final_error": err,
                   "rows": rows, "weights": w.to_literals()}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
