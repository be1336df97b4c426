"""TEST INFRASTRUCTURE ONLY.  ctypes loaders for the CPU oracles plus a float64 NumPy restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

  port()  -> oracle/liboracle.so      : C restatement of CE_Predictive_Node.cpp:146-261 (crn_oracle.c)
  ref()   -> oracle/_ref/libcrn_ref.so: the reference's UNMODIFIED engine object (ref_harness.cpp), or
                                        None when it has not been built (no /root/reference)
  sense_f64(): float64 NumPy (pocketfft) restatement, an independent bound on both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libcrn_ref.so")

_port = None
_ref = None


def build():
    """Compile the C restatement and, when the reference tree is present, oracle O1."""
    subprocess.run(["make", "-s", "-C", HERE], check=True, capture_output=True)


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            build()
        lib = C.CDLL(PORT_PATH)
        lib.crn_oracle_sense.restype = C.c_int
        lib.crn_oracle_sense.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int]
        lib.crn_oracle_time.restype = C.c_double
        lib.crn_oracle_time.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        lib.crn_oracle_max_threads.restype = C.c_int
        lib.crn_oracle_hann.restype = C.c_float
        lib.crn_oracle_hann.argtypes = [C.c_int, C.c_int]
        lib.crn_oracle_stream_seed.restype = C.c_uint64
        lib.crn_oracle_stream_seed.argtypes = [C.c_uint64, C.c_int64]
        lib.crn_oracle_pu_states.restype = None
        lib.crn_oracle_pu_states.argtypes = [C.c_uint64, C.c_int, C.c_int64, C.c_void_p]
        lib.crn_oracle_pu_next.restype = C.c_int
        lib.crn_oracle_pu_next.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.crn_oracle_synth.restype = None
        lib.crn_oracle_synth.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        lib.crn_oracle_synth_sigma2.restype = C.c_double
        lib.crn_oracle_synth_sigma2.argtypes = [C.c_void_p]
        lib.crn_oracle_ann_forward.restype = None
        lib.crn_oracle_ann_forward.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        lib.crn_oracle_ann_error.restype = C.c_double
        lib.crn_oracle_ann_error.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        lib.crn_oracle_ann_train.restype = C.c_int
        lib.crn_oracle_ann_train.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        _port = lib
    return _port


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_PATH):
            return None
        lib = C.CDLL(REF_PATH)
        lib.crn_ref_run.restype = C.c_long
        lib.crn_ref_run.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_long]
        lib.crn_ref_time.restype = C.c_double
        lib.crn_ref_time.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_int, C.POINTER(C.c_long)]
        lib.crn_ref_constants.restype = None
        lib.crn_ref_constants.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _ref = lib
    return _ref


def _iq_f32(iq):
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    return iq


def sense_port(cfg, iq, ngroups=None, nthreads=1):
    """cfg: crn_b200.Config (same struct the GPU library takes).  Returns (feat, ann, decision, mask)."""
    stride = cfg.frame_stride if cfg.frame_stride > 0 else cfg.frame_len
    gs = stride * cfg.navg
    if getattr(cfg, "iq_format", 0) == 1:      # sc16: interleaved int16 (I,Q)
        iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
        nsamp = iq.size // 2
    else:
        iq = _iq_f32(iq)
        nsamp = iq.size
    if ngroups is None:
        ngroups = nsamp // gs
    assert ngroups * gs <= nsamp
    feat = np.zeros((ngroups, cfg.nbands), np.float32)
    ann = np.zeros((ngroups, 3), np.float64)
    dec = np.zeros(ngroups, np.int32)
    mask = np.zeros(ngroups, np.uint64)
    rc = port().crn_oracle_sense(C.byref(cfg), iq.ctypes.data, ngroups, feat.ctypes.data, ann.ctypes.data,
                                 dec.ctypes.data, mask.ctypes.data, nthreads)
    assert rc == 0
    return feat, ann, dec, mask


def time_port(cfg, iq, ngroups, nthreads):
    iq = np.ascontiguousarray(iq, dtype=np.int16) if getattr(cfg, "iq_format", 0) == 1 else _iq_f32(iq)
    return port().crn_oracle_time(C.byref(cfg), iq.ctypes.data, ngroups, nthreads)


def sense_ref(iq, L=512, want_bins=False):
    """Oracle O1: the unmodified reference engine on frames of L <= 512 samples.
    Returns (feat[nd,4] = NF^2,CH1,CH2,CH3, ann[nd,3], decision[nd], tx_freq[nd][, avg_bins[nd,512]])."""
    lib = ref()
    assert lib is not None, "oracle/_ref/libcrn_ref.so not built"
    iq = _iq_f32(iq)
    nframes = iq.size // L
    nd = nframes // 10
    feat = np.zeros((nd, 4), np.float32)
    ann = np.zeros((nd, 3), np.float64)
    dec = np.zeros(nd, np.int32)
    txf = np.zeros(nd, np.float64)
    bins = np.zeros((nd, 512), np.float32) if want_bins else None
    got = lib.crn_ref_run(iq.ctypes.data, L, nframes, feat.ctypes.data, ann.ctypes.data, dec.ctypes.data,
                          txf.ctypes.data, bins.ctypes.data if want_bins else None, nd)
    assert got == nd, (got, nd)
    return (feat, ann, dec, txf, bins) if want_bins else (feat, ann, dec, txf)


def time_ref(iq, L, nframes, nthreads):
    iq = _iq_f32(iq)
    nd = C.c_long(0)
    sec = ref().crn_ref_time(iq.ctypes.data, L, nframes, nthreads, C.byref(nd))
    return sec, nd.value


def hann(N):
    return np.array([port().crn_oracle_hann(n, N) for n in range(N)], np.float32)


def sense_f64(cfg, iq, ngroups=None):
    """Oracle O2: float64 restatement (NumPy pocketfft) of CE_Predictive_Node.cpp:146-261 with the same
    options.  Independent of the C FFT restatement; bounds both the port and the GPU path."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    N, L, K = cfg.nfft, cfg.frame_len, cfg.navg
    stride = cfg.frame_stride if cfg.frame_stride > 0 else L
    gs = stride * K
    if ngroups is None:
        ngroups = iq.size // gs
    fr = iq[: ngroups * gs].reshape(ngroups, K, stride)[:, :, :L].astype(np.complex128)
    if cfg.window == 1:
        n = np.arange(N)
        w = 0.5 - 0.5 * np.cos(2 * np.pi * n / (N - 1))
        fr = fr * w[:L]
    X = np.fft.fft(fr, n=N, axis=-1)
    d = np.abs(X) if cfg.detector == 0 else (X.real ** 2 + X.imag ** 2)
    avg = d.sum(axis=1) / K
    m = np.zeros((ngroups, cfg.nbands))
    for s in range(cfg.nsegs):
        sg = cfg.segs[s]
        m[:, sg.band] += avg[:, sg.lo:sg.hi].sum(axis=1)
    feat = m * m if cfg.postop == 0 else m
    ann = np.zeros((ngroups, 3))
    dec = np.zeros(ngroups, np.int32)
    if cfg.decide == 1:
        ann, dec = mlp_f64(cfg, feat)
    return feat, ann, dec


def mlp_f64(cfg, feat):
    """CE_Predictive_Node.cpp:200,214-261 on given features (rows = NF^2, CH1, CH2, CH3), float64."""
    feat = np.asarray(feat, np.float64)
    n = feat.shape[0]
    wih = np.array([[cfg.ann_wih[i][j] for j in range(6)] for i in range(5)])
    who = np.array([[cfg.ann_who[j][k] for k in range(4)] for j in range(6)])
    F = np.concatenate([np.ones((n, 1)), feat[:, :4]], axis=1)  # bias at index 0
    with np.errstate(over="ignore"):
        H = 1.0 / (1.0 + np.exp(-(F @ wih[:, 1:])))
        Hb = np.concatenate([np.ones((n, 1)), H], axis=1)
        ann = 1.0 / (1.0 + np.exp(-(Hb @ who[:, 1:])))
    thr = cfg.ann_threshold
    dec = np.where(ann[:, 0] >= thr, 1, np.where(ann[:, 1] >= thr, 2, np.where(ann[:, 2] >= thr, 3, 0))).astype(np.int32)
    return ann, dec


def synth(sc, nsamples, first=0, stream=0):
    """CPU statement of the synthetic PU capture (crn_oracle_synth).  Returns (iq complex64, states int8)."""
    lib = port()
    dwell = sc.dwell_groups * sc.group_samples
    ndwell = (first + nsamples + dwell - 1) // dwell
    sseed = lib.crn_oracle_stream_seed(sc.seed, stream)
    states = np.zeros(ndwell, np.int8)
    lib.crn_oracle_pu_states(sseed, sc.hop_mode, ndwell, states.ctypes.data)
    iq = np.zeros(nsamples, np.complex64)
    lib.crn_oracle_synth(C.byref(sc), sseed, states.ctypes.data, iq.ctypes.data, first, nsamples)
    return iq, states


# ---- the occupancy predictor on its own (checker for crn_ann_forward_device / crn_ann_train_device) -------------

def ann_forward(weights, feat, threshold=0.8):
    """CE_Predictive_Node.cpp:200,214-261 on feature rows (float32 [n, >=4]).  Returns (out[n,3] f64, decision[n])."""
    feat = np.ascontiguousarray(feat, np.float32)
    n, stride = feat.shape
    out = np.zeros((n, 3), np.float64)
    dec = np.zeros(n, np.int32)
    port().crn_oracle_ann_forward(C.byref(weights), threshold, feat.ctypes.data, n, stride, out.ctypes.data, dec.ctypes.data)
    return out, dec


def ann_error(weights, scale, feat, labels):
    """E = 1/2 sum (t - Output)^2 and the summed descent direction -dE/dW at (scaled-input) weights."""
    feat = np.ascontiguousarray(feat, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    scale = np.ascontiguousarray(scale, np.float64)
    g = type(weights)()
    E = port().crn_oracle_ann_error(C.byref(weights), scale.ctypes.data, feat.ctypes.data, feat.shape[1],
                                    labels.ctypes.data, feat.shape[0], C.byref(g))
    return E, g


def ann_train(tc, feat, labels, weights):
    """Serial batch back-propagation (the CPU statement of crn_ann_train_device).  Returns (weights, E, epochs)."""
    feat = np.ascontiguousarray(feat, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    w = type(weights)()
    C.memmove(C.byref(w), C.byref(weights), C.sizeof(w))
    err, ep = C.c_double(0.0), C.c_int32(0)
    rc = port().crn_oracle_ann_train(C.byref(tc), feat.ctypes.data, feat.shape[1], labels.ctypes.data, feat.shape[0],
                                     C.byref(w), C.byref(err), C.byref(ep))
    assert rc == 0
    return w, err.value, ep.value
