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
_native = None

# ---- the oracle's own mirror of include/crnsense.h's crn_config / crn_synth_config (layout only): the reference
# arm of bench.py builds its workloads from these and crn_oracle_config.c, without loading the product library.
MAX_BANDS, MAX_SEGS = 64, 128


class Seg(C.Structure):
    _fields_ = [("band", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32)]


class Config(C.Structure):
    _fields_ = [
        ("nfft", C.c_int32), ("frame_len", C.c_int32), ("frame_stride", C.c_int32), ("navg", C.c_int32),
        ("window", C.c_int32), ("detector", C.c_int32), ("postop", C.c_int32), ("decide", C.c_int32),
        ("nbands", C.c_int32), ("nsegs", C.c_int32), ("segs", Seg * MAX_SEGS),
        ("ann_wih", (C.c_double * 6) * 5), ("ann_who", (C.c_double * 4) * 6),
        ("ann_threshold", C.c_double), ("energy_factor", C.c_double),
        ("device", C.c_int32), ("ring_slots", C.c_int32), ("iq_format", C.c_int32), ("reserved_", C.c_int32),
    ]

    @property
    def group_samples(self):
        return (self.frame_stride if self.frame_stride > 0 else self.frame_len) * self.navg


class SynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("fs", C.c_double), ("pu_rate", C.c_double), ("offsets_hz", C.c_double * 3),
        ("snr_db", C.c_double), ("pu_gain_db", C.c_double), ("hop_mode", C.c_int32),
        ("dwell_groups", C.c_int32), ("group_samples", C.c_int32),
        ("intf_type", C.c_int32), ("intf_period_groups", C.c_int32), ("pu_framed", C.c_int32),
        ("intf_offset_hz", C.c_double), ("intf_rate", C.c_double), ("intf_gain_db", C.c_double),
        ("intf_duty", C.c_double),
    ]


def config_reference():
    c = Config()
    port().crn_oracle_config_reference(C.byref(c))
    return c


def config_welch(nfft, navg):
    c = Config()
    assert port().crn_oracle_config_welch(C.byref(c), nfft, navg) == 0, (nfft, navg)
    return c


def config_wideband(nfft, navg, nch):
    c = Config()
    assert port().crn_oracle_config_wideband(C.byref(c), nfft, navg, nch) == 0, (nfft, navg, nch)
    return c


def synth_config(group_samples, **kw):
    sc = SynthConfig()
    port().crn_oracle_synth_config_default(C.byref(sc), group_samples)
    for k, v in kw.items():
        if k == "offsets_hz":
            for i, x in enumerate(v):
                sc.offsets_hz[i] = x
        else:
            assert hasattr(sc, k), k
            setattr(sc, k, v)
    return sc


def build():
    """Compile the C restatement and, when the reference tree is present, oracle O1."""
    subprocess.run(["make", "-s", "-C", HERE], check=True, capture_output=True)


def _declare(lib):
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
    lib.crn_oracle_config_reference.restype = None
    lib.crn_oracle_config_reference.argtypes = [C.c_void_p]
    lib.crn_oracle_config_welch.restype = C.c_int
    lib.crn_oracle_config_welch.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.crn_oracle_config_wideband.restype = C.c_int
    lib.crn_oracle_config_wideband.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.crn_oracle_synth_config_default.restype = None
    lib.crn_oracle_synth_config_default.argtypes = [C.c_void_p, C.c_int32]
    return lib


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            build()
        _port = _declare(C.CDLL(PORT_PATH))
    return _port


def native():
    """The port compiled -O3 -march=native ON THIS HOST (timing only; bench.py's cpu_baseline).  None if it cannot
    be built here."""
    global _native
    if _native is None:
        out = os.path.join("/tmp", "crn_oracle_native_%d" % os.getuid())
        try:
            subprocess.run(["make", "-s", "-C", HERE, "native", "NATIVE_OUT=" + out], check=True, capture_output=True)
            _native = _declare(C.CDLL(os.path.join(out, "liboracle_native.so")))
        except (OSError, subprocess.CalledProcessError):
            _native = False
    return _native or None


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


def time_port(cfg, iq, ngroups, nthreads, lib=None):
    iq = np.ascontiguousarray(iq, dtype=np.int16) if getattr(cfg, "iq_format", 0) == 1 else _iq_f32(iq)
    return (lib or port()).crn_oracle_time(C.byref(cfg), iq.ctypes.data, ngroups, nthreads)


def fftw_installed():
    """Is an FFTW3 single-precision library on this host?  (liquid-dsp can be configured on top of it; the restated
    radix-2 FFT of the port is what liquid does without it.)  Reported in cpu_baseline.sample, nothing links it."""
    import ctypes.util
    return ctypes.util.find_library("fftw3f") is not None


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
