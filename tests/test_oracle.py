"""The oracle is pinned before it is trusted: port == unmodified reference engine (bit exact), port ==
golden fixtures produced by the reference engine, known answers derived from the reference's literals,
analytic DFT facts, float64 bound."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, feat_close


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


def test_golden_fixtures_exist():
    assert len(golden_cases()) >= 5


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_port_reproduces_reference_engine_fixtures_bit_exact(crn, oracle, path):
    g = np.load(path)
    cfg = crn.config_reference()
    cfg.frame_len = int(g["L"])
    feat, ann, dec, _ = oracle.sense_port(cfg, g["iq"])
    assert np.array_equal(feat, g["feat"])          # fp32, same operations in the same order
    assert np.array_equal(ann, g["ann"])
    assert np.array_equal(dec, g["decision"])
    # the retune the reference performed (CE_Predictive_Node.cpp:245-261)
    want_tx = np.array([crn.TX_FREQ_FOR_DECISION[int(d)] or 0.0 for d in dec])
    assert np.array_equal(want_tx, g["tx_freq"])


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_float64_restatement_bounds_the_fixtures(crn, oracle, path):
    g = np.load(path)
    cfg = crn.config_reference()
    cfg.frame_len = int(g["L"])
    feat, ann, dec = oracle.sense_f64(cfg, g["iq"])
    assert feat_close(feat, g["feat"], 2e-5)
    assert np.abs(ann - g["ann"]).max() <= 1e-6
    assert np.array_equal(dec, g["decision"])


def test_port_equals_live_reference_engine(crn, oracle):
    """When oracle/_ref is built (build container, or shipped prebuilt) run the reference engine itself."""
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built here")
    import ctypes as C
    n, k = C.c_int(), C.c_int()
    oracle.ref().crn_ref_constants(C.byref(n), C.byref(k))
    assert (n.value, k.value) == (512, 10)  # CE_Predictive_Node.hpp:31-32
    for L, seed, snr in ((512, 21, 10.0), (363, 22, 0.0), (1, 23, 20.0), (511, 24, 20.0)):
        gs = L * 10
        sc = crn.synth_config(gs, dwell_groups=1, snr_db=snr, seed=seed)
        iq, _ = oracle.synth(sc, 12 * gs)
        cfg = crn.config_reference()
        cfg.frame_len = L
        f0, a0, d0, _, bins = oracle.sense_ref(iq, L=L, want_bins=True)
        f1, a1, d1, _ = oracle.sense_port(cfg, iq)
        assert np.array_equal(f0, f1) and np.array_equal(a0, a1) and np.array_equal(d0, d1)
        # the engine's averaged spectrum is non-negative and its band sums give the features
        assert (bins >= 0).all()
        m2 = bins[:, 55:85].astype(np.float32).sum(axis=1, dtype=np.float32)
        assert np.allclose(m2 * m2, f0[:, 2], rtol=1e-5)


def test_ann_known_answers_from_reference_weights(crn, oracle):
    """outputs implied by the 43 literals of CE_Predictive_Node.cpp:78-120."""
    g = np.load(os.path.join(GOLDEN, "ref_zeros.npz"))
    assert np.allclose(g["ann"][0], [4.78996574e-01, 4.12229629e-05, 3.35047425e-03], rtol=1e-7)
    assert g["decision"][0] == 0 and g["tx_freq"][0] == 0.0  # ALL BUSY: no retune
    assert np.array_equal(g["feat"][0], np.zeros(4, np.float32))

    # drive the MLP alone through the port: one bin per band carries sqrt(feature)
    def mlp(nf, c1, c2, c3):
        cfg = crn.config_reference()
        cfg.navg = 1
        X = np.zeros(512, np.complex128)
        for b, v in ((300, nf), (0, c1), (55, c2), (189, c3)):
            X[b] = np.sqrt(v)
        x = np.fft.ifft(X).astype(np.complex64)
        feat, ann, dec, _ = oracle.sense_port(cfg, x)
        assert np.allclose(feat[0], [nf, c1, c2, c3], rtol=1e-4)
        return ann[0], int(dec[0])
    out, dec = mlp(1e4, 1e8, 1e5, 1e5)
    assert dec == 1 and abs(out[0] - 0.99943) < 1e-4 and out[1] < 1e-6
    out, dec = mlp(1e4, 1e5, 1e8, 1e5)
    assert dec == 2 and abs(out[1] - 0.99941) < 1e-4
    out, dec = mlp(1e4, 1e5, 1e5, 1e8)
    assert dec == 3 and abs(out[2] - 0.99947) < 1e-4


def test_tone_fixture_is_analytic(crn):
    """One unit tone at bin 70 for L = N = 512: |X[70]| = 512, every other bin ~0, K-average keeps it:
    CH2 = 512^2, the other features ~0 (CE_Predictive_Node.cpp:152-154,181-183,195)."""
    g = np.load(os.path.join(GOLDEN, "ref_tone70.npz"))
    assert abs(g["avg_bins"][0, 70] - 512.0) < 1e-2
    assert np.delete(g["avg_bins"][0], 70).max() < 1e-2
    assert abs(g["feat"][0, 2] - 512.0 ** 2) / 512.0 ** 2 < 1e-5
    assert g["decision"][0] == 2 and g["tx_freq"][0] == 833e6


@pytest.mark.parametrize("nfft,navg,mode", [(256, 3, "ref"), (1024, 64, "welch"), (2048, 4, "welch"),
                                             (8192, 2, "wide"), (4096, 1, "wide")])
def test_port_vs_float64_extension_modes(crn, oracle, nfft, navg, mode):
    if mode == "ref":
        cfg = crn.config_reference()
        cfg.nfft, cfg.frame_len, cfg.navg = nfft, nfft, navg
        for i in range(cfg.nsegs):
            cfg.segs[i].lo //= 2
            cfg.segs[i].hi //= 2
    elif mode == "welch":
        cfg = crn.config_welch(nfft, navg)
    else:
        cfg = crn.config_wideband(nfft, navg, 64)
    gs = cfg.group_samples
    sc = crn.synth_config(gs, dwell_groups=1, snr_db=5.0, seed=nfft)
    iq, _ = oracle.synth(sc, 2 * gs)
    f1, a1, d1, m1 = oracle.sense_port(cfg, iq)
    f2, a2, d2 = oracle.sense_f64(cfg, iq)
    assert feat_close(f1, f2, 2e-5)
    if cfg.decide == crn.DECIDE_ANN:
        assert np.abs(a1 - a2).max() < 1e-6 and np.array_equal(d1, d2)


def test_port_threads_agree(crn, oracle):
    cfg = crn.config_welch(1024, 8)
    sc = crn.synth_config(cfg.group_samples, dwell_groups=1)
    iq, _ = oracle.synth(sc, 9 * cfg.group_samples)
    a = oracle.sense_port(cfg, iq, nthreads=1)
    b = oracle.sense_port(cfg, iq, nthreads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_pu_hop_models(oracle):
    """CE_PU_MARKOV_Chain_Tx.cpp:97-128 as coded vs README.md:70-74 as documented vs CE_Random_Behaviour_PU."""
    nxt = oracle.port().crn_oracle_pu_next
    for cur in range(3):
        assert nxt(1, cur, 0) == 0 and all(nxt(1, cur, r) == 1 for r in range(1, 10))  # as coded: never CH3
        assert nxt(0, cur, 0) == 0
    assert [nxt(0, 0, r) for r in range(10)] == [0, 1, 1, 1, 2, 2, 2, 2, 2, 2]   # .1 .3 .6
    assert [nxt(0, 1, r) for r in range(10)] == [0, 1, 1, 1, 1, 1, 2, 2, 2, 2]   # .1 .5 .4
    assert [nxt(0, 2, r) for r in range(10)] == [0, 1, 1, 2, 2, 2, 2, 2, 2, 2]   # .1 .2 .7
    assert [nxt(2, 0, r) for r in range(3)] == [0, 1, 2]
    st = np.zeros(20000, np.int8)
    oracle.port().crn_oracle_pu_states(oracle.port().crn_oracle_stream_seed(12, 0), 0, st.size, st.ctypes.data)
    assert st[0] == 0
    # empirical transition rows of the documented chain
    for cur, row in ((0, [.1, .3, .6]), (1, [.1, .5, .4]), (2, [.1, .2, .7])):
        idx = np.where(st[:-1] == cur)[0]
        emp = np.bincount(st[idx + 1], minlength=3) / len(idx)
        assert np.abs(emp - row).max() < 0.03


def test_synth_statistics(crn, oracle):
    """Unit-power OFDM x 10^(-12/20) at the chosen offset plus AWGN at the stated in-band SNR."""
    gs = 65536
    sc = crn.synth_config(gs, dwell_groups=4, snr_db=10.0, seed=5)
    iq, states = oracle.synth(sc, 4 * gs)
    assert states[0] == 0
    sig2 = oracle.port().crn_oracle_synth_sigma2(__import__("ctypes").byref(sc))
    p = np.mean(np.abs(iq) ** 2)
    assert abs(p - (10 ** -1.2 + sig2)) / p < 0.05
    # spectrum: PU energy sits within +-0.56 MHz of the CH1 offset (0 Hz)
    X = np.abs(np.fft.fft(iq[: 4096 * 32].reshape(32, 4096), axis=1)) ** 2
    psd = X.mean(axis=0)
    hz = np.fft.fftfreq(4096, 1 / 13e6)
    inband = psd[np.abs(hz) < 0.5e6].mean()
    outband = psd[np.abs(hz) > 1.0e6].mean()
    assert 8.0 < 10 * np.log10(inband / outband) < 13.0   # (S+N)/N at 10 dB in-band SNR = 10.4 dB
    # deterministic and position independent
    iq2, _ = oracle.synth(sc, 1000, first=gs + 17)
    assert np.array_equal(iq2, iq[gs + 17: gs + 1017])


def test_synth_interferer_waveforms(crn, oracle):
    """The interferer node's modem-free waveforms (src/interferer.cpp:128-154) on top of the PU capture: what is
    added is exactly the documented waveform - checked by subtracting the interferer-free capture."""
    gs = 4096
    base_cfg = dict(dwell_groups=4, snr_db=10.0, seed=9)
    clean, _ = oracle.synth(crn.synth_config(gs, **base_cfg), 8 * gs)
    off, g = -5.2e6, 10 ** (-3.0 / 20.0)
    n = np.arange(8 * gs)
    carrier = np.exp(2j * np.pi * ((n * (off / 13e6)) % 1.0))
    # CW: the constant 0.5 + 0.5j (BuildCWTransmission) at the interferer's offset, duty cycle 1/2 of 4 groups
    sc = crn.synth_config(gs, intf_type=crn.INTF_CW, intf_offset_hz=off, intf_period_groups=4, intf_duty=0.5, **base_cfg)
    d = oracle.synth(sc, 8 * gs)[0] - clean
    on = ((n // gs) % 4) < 2
    assert np.abs(d[~on]).max() == 0.0
    assert np.abs(d[on] - (g * (0.5 + 0.5j) * carrier)[on]).max() <= 2e-5
    # NOISE: uniform in [-0.25, 0.25) per component, one draw per interferer sample (1 MS/s held to 13 MS/s)
    sc = crn.synth_config(gs, intf_type=crn.INTF_NOISE, intf_offset_hz=off, **base_cfg)
    b = (oracle.synth(sc, 8 * gs)[0] - clean) / (g * carrier)
    assert np.abs(b.real).max() <= 0.2501 and np.abs(b.imag).max() <= 0.2501
    assert abs(b.real.mean()) < 0.01 and abs(b.real.var() - 0.25 / 12) < 0.003
    held = (n * (1e6 / 13e6)).astype(np.int64)
    same = held[1:] == held[:-1]
    assert np.abs(b[1:][same] - b[:-1][same]).max() <= 2e-5 and np.abs(b[1:][~same] - b[:-1][~same]).mean() > 0.05
    # AWGN as coded: normal_distribution(5.0, 5.0) per component -> a DC offset of 5 + 5j rides on the noise
    sc = crn.synth_config(gs, intf_type=crn.INTF_AWGN, intf_offset_hz=off, intf_rate=13e6, **base_cfg)
    b = (oracle.synth(sc, 8 * gs)[0] - clean) / (g * carrier)
    assert abs(b.real.mean() - 5.0) < 0.15 and abs(b.imag.mean() - 5.0) < 0.15
    assert abs(b.real.std() - 5.0) < 0.15 and abs(b.imag.std() - 5.0) < 0.15
    # and the sensing path sees it: a CW in the noise-floor band (bins 300..309 of 512 = -5.38..-5.15 MHz) lifts NF^2
    cfg = crn.config_reference()
    cap = crn.synth_config(cfg.group_samples, dwell_groups=2, snr_db=10.0, seed=3)
    jam = crn.synth_config(cfg.group_samples, dwell_groups=2, snr_db=10.0, seed=3, intf_type=crn.INTF_CW,
                           intf_offset_hz=-5.25e6)
    f0 = oracle.sense_port(cfg, oracle.synth(cap, 4 * cfg.group_samples)[0])[0]
    f1 = oracle.sense_port(cfg, oracle.synth(jam, 4 * cfg.group_samples)[0])[0]
    assert (f1[:, 0] > 50 * f0[:, 0]).all() and np.allclose(f1[:, 1:], f0[:, 1:], rtol=0.05)


def _cfg_fields(c):
    return dict(nfft=c.nfft, frame_len=c.frame_len, frame_stride=c.frame_stride, navg=c.navg, window=c.window,
                detector=c.detector, postop=c.postop, decide=c.decide, nbands=c.nbands, nsegs=c.nsegs,
                segs=[(c.segs[i].band, c.segs[i].lo, c.segs[i].hi) for i in range(c.nsegs)],
                wih=[[c.ann_wih[i][j] for j in range(6)] for i in range(5)],
                who=[[c.ann_who[j][k] for k in range(4)] for j in range(6)],
                thr=c.ann_threshold, ef=c.energy_factor, ring=c.ring_slots, fmt=c.iq_format)


def test_oracle_configs_equal_the_product_fillers(crn, oracle):
    """The oracle states the workloads on its own (oracle/crn_oracle_config.c, restated from the reference's literals
    and loop bounds) so that bench.py's reference arm never loads the product library; the two statements must say
    the same thing, field for field, and the struct mirrors must have the same layout."""
    import ctypes as C
    assert C.sizeof(oracle.Config) == C.sizeof(crn.Config) and C.sizeof(oracle.SynthConfig) == C.sizeof(crn.SynthConfig)
    for name, _ in crn.Config._fields_:
        assert getattr(oracle.Config, name).offset == getattr(crn.Config, name).offset, name
    for name, _ in crn.SynthConfig._fields_:
        assert getattr(oracle.SynthConfig, name).offset == getattr(crn.SynthConfig, name).offset, name
    assert _cfg_fields(oracle.config_reference()) == _cfg_fields(crn.config_reference())
    for nfft in (256, 512, 1024, 2048, 4096, 8192):
        assert _cfg_fields(oracle.config_welch(nfft, 64)) == _cfg_fields(crn.config_welch(nfft, 64)), nfft
        nch = 16 if nfft == 256 else 64
        assert _cfg_fields(oracle.config_wideband(nfft, 7, nch)) == _cfg_fields(crn.config_wideband(nfft, 7, nch)), nfft
    a, b = oracle.synth_config(65536, snr_db=3.0, hop_mode=2), crn.synth_config(65536, snr_db=3.0, hop_mode=2)
    assert bytes(a) == bytes(b)


def test_welch_band_plan_at_256_points(crn):
    """N = 256 halves the reference's bin indices (bins are twice as wide); every segment stays non-empty, ordered and
    inside the spectrum, and bin N-1 is still excluded from CH1 as upstream excludes 511 (.cpp:177)."""
    c = crn.config_welch(256, 10)
    assert crn.validate(c) == crn.OK
    segs = [(c.segs[i].band, c.segs[i].lo, c.segs[i].hi) for i in range(c.nsegs)]
    assert segs == [(0, 150, 155), (1, 0, 8), (1, 248, 255), (2, 27, 42), (3, 94, 111)]


def test_reference_arm_never_loads_the_product_library():
    """bench.py --impl reference in a fresh interpreter: its JSON line lists the shared objects of this repo it mapped."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["native_so_mapped"] and all(m.startswith("oracle/") for m in line["native_so_mapped"]), line["native_so_mapped"]


def test_modem_waveforms_have_their_spectra(oracle):
    """CPU statement of the interferer modems (src/interferer.cpp:156-288), observed alone at their own rate: GMSK is
    constant-envelope with ~1.04 Rs of 99 % bandwidth (BT = 0.5, 4 samples/symbol), root-raised-cosine QPSK stays
    inside (1 + beta) Rs, the OFDM burst fills 51/64 of the band with a Gaussian-like envelope."""
    gs = 65536

    def facts(itype):
        sc = oracle.synth_config(gs, pu_gain_db=-300.0, snr_db=200.0, intf_type=itype, intf_rate=13e6, intf_gain_db=0.0)
        iq, _ = oracle.synth(sc, 4 * gs)
        p = np.abs(iq) ** 2
        X = np.fft.fftshift(np.abs(np.fft.fft(iq.reshape(-1, 4096), axis=1)) ** 2).mean(axis=0)
        c = np.cumsum(X) / X.sum()
        return p.mean(), 10 * np.log10(p.max() / p.mean()), (np.searchsorted(c, 0.995) - np.searchsorted(c, 0.005)) / 4096.0

    pw, papr, bw = facts(4)
    assert abs(pw - 1.0) < 0.01 and papr < 0.05 and 0.24 < bw < 0.28
    pw, papr, bw = facts(5)
    assert 0.07 < pw < 0.10 and 3.0 < papr < 7.0 and 0.5 < bw < 0.675
    pw, papr, bw = facts(6)
    assert abs(pw - 1.0) < 0.05 and 8.0 < papr < 13.0 and 0.76 < bw < 0.83
