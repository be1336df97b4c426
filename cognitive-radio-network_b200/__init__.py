"""crn_b200 — Python host side of libcrnsense (ctypes over the C-ABI in include/crnsense.h).

This is plumbing for tests and bench.py: the product is the CUDA library.  Nothing in this package
imports, loads or calls anything under oracle/, and there is no CPU fallback: if libcrnsense.so is
missing the import fails loudly, and without a CUDA device `Sensor()` raises.

Names mirror the reference's engine (cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.{hpp,cpp}):
a `Sensor` is one CE_Predictive_Node sensing state (one stream on one GPU); `push_frame` is what
execute() does per USRP_RX_SAMPS event (.cpp:146-154); a finished `Result` carries what the reference
only printf()s (.cpp:163-261).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRN_LIB") or os.path.join(_HERE, "libcrnsense.so")  # CRN_LIB: A/B build variants

MAX_BANDS = 64
MAX_SEGS = 128

OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_OVERRUN, ERR_NOT_READY, ERR_UNSUPPORTED = range(-1, -8, -1)
WINDOW_RECT, WINDOW_HANN = 0, 1
DET_MAG, DET_MAGSQ = 0, 1
POST_SQUARE_OF_SUM, POST_SUM, POST_SUM_DB = 0, 1, 2
FUSE_OR, FUSE_MAJORITY, FUSE_AND = 0, 1, 2
DECIDE_NONE, DECIDE_ANN, DECIDE_ENERGY = 0, 1, 2
ALL_BUSY, CH1_OCCUPIED, CH2_OCCUPIED, CH3_OCCUPIED = 0, 1, 2, 3
IQ_CF32, IQ_SC16 = 0, 1
INTF_NONE, INTF_CW, INTF_NOISE, INTF_AWGN = 0, 1, 2, 3  # src/interferer.cpp waveforms that need no modem
INTF_GMSK, INTF_RRC, INTF_OFDM = 4, 5, 6              # ... and the framed ones (:156-288)

# TX retune the reference performs for each decision (CE_Predictive_Node.cpp:245-261, .hpp:55-57)
TX_FREQ_FOR_DECISION = {ALL_BUSY: None, CH1_OCCUPIED: 835e6, CH2_OCCUPIED: 833e6, CH3_OCCUPIED: 835e6}


class Seg(C.Structure):
    _fields_ = [("band", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32)]


class Config(C.Structure):
    _fields_ = [
        ("nfft", C.c_int32), ("frame_len", C.c_int32), ("frame_stride", C.c_int32), ("navg", C.c_int32),
        ("window", C.c_int32), ("detector", C.c_int32), ("postop", C.c_int32), ("decide", C.c_int32),
        ("nbands", C.c_int32), ("nsegs", C.c_int32),
        ("segs", Seg * MAX_SEGS),
        ("ann_wih", (C.c_double * 6) * 5),
        ("ann_who", (C.c_double * 4) * 6),
        ("ann_threshold", C.c_double), ("energy_factor", C.c_double),
        ("device", C.c_int32), ("ring_slots", C.c_int32),
        ("iq_format", C.c_int32), ("reserved_", C.c_int32),
    ]

    def copy(self):
        c = Config()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Config))
        return c

    @property
    def stride(self):
        return self.frame_stride if self.frame_stride > 0 else self.frame_len

    @property
    def group_samples(self):
        return self.stride * self.navg


class Result(C.Structure):
    _fields_ = [
        ("first_frame", C.c_uint64), ("decision", C.c_int32), ("nfeat", C.c_int32),
        ("occupancy_mask", C.c_uint64), ("ann_out", C.c_double * 3), ("feat", C.c_float * MAX_BANDS),
    ]


class SynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("fs", C.c_double), ("pu_rate", C.c_double), ("offsets_hz", C.c_double * 3),
        ("snr_db", C.c_double), ("pu_gain_db", C.c_double), ("hop_mode", C.c_int32),
        ("dwell_groups", C.c_int32), ("group_samples", C.c_int32),
        ("intf_type", C.c_int32), ("intf_period_groups", C.c_int32), ("pu_framed", C.c_int32),
        ("intf_offset_hz", C.c_double), ("intf_rate", C.c_double), ("intf_gain_db", C.c_double),
        ("intf_duty", C.c_double),
    ]


class AnnWeights(C.Structure):
    """WeightIH / WeightHO with the reference's 1-based indexing (CE_Predictive_Node.hpp:66-72)."""
    _fields_ = [("wih", (C.c_double * 6) * 5), ("who", (C.c_double * 4) * 6)]

    @classmethod
    def from_config(cls, cfg):
        w = cls()
        C.memmove(C.byref(w), C.addressof(cfg) + Config.ann_wih.offset, C.sizeof(cls))
        return w

    def into_config(self, cfg):
        C.memmove(C.addressof(cfg) + Config.ann_wih.offset, C.byref(self), C.sizeof(AnnWeights))
        return cfg

    def arrays(self):
        return (np.array([[self.wih[i][j] for j in range(6)] for i in range(5)]),
                np.array([[self.who[j][k] for k in range(4)] for j in range(6)]))

    def to_literals(self):
        """The weights written like the reference's assignment block (CE_Predictive_Node.cpp:78-120); this is the
        file format of the engine's `-m <file>` ce_args option (%.17g: the doubles round-trip exactly)."""
        lines = ["WeightIH[%d][%d]   =        %.17g;" % (i, j, self.wih[i][j]) for j in range(1, 6) for i in range(5)]
        lines += ["WeightHO[%d][%d]   =        %.17g;" % (j, k, self.who[j][k]) for k in range(1, 4) for j in range(6)]
        return "\n".join(lines) + "\n"

    @classmethod
    def from_literals(cls, text):
        import re
        w = cls()
        for name, a, b, v in re.findall(r"Weight(IH|HO)\[(\d+)\]\[(\d+)\]\s*=\s*([-+0-9.eE]+)\s*;", text):
            (w.wih if name == "IH" else w.who)[int(a)][int(b)] = float(v)
        return w


class AnnTrainConfig(C.Structure):
    _fields_ = [
        ("max_epochs", C.c_int32), ("check_every", C.c_int32), ("eta", C.c_double), ("alpha", C.c_double),
        ("target_error", C.c_double), ("input_scale", C.c_double * 4), ("init_range", C.c_double),
        ("seed", C.c_uint64),
    ]


class KernelInfo(C.Structure):
    _fields_ = [
        ("nfft", C.c_int32), ("threads_per_frame", C.c_int32), ("elems_per_thread", C.c_int32),
        ("teams_per_cta", C.c_int32), ("threads_per_cta", C.c_int32), ("ctas_per_sm", C.c_int32),
        ("grid", C.c_int32), ("smem_bytes", C.c_int32), ("regs_per_thread", C.c_int32),
        ("num_sms", C.c_int32), ("name", C.c_char * 64),
    ]


# every symbol include/crnsense.h declares: (restype, argtypes)
_P = C.c_void_p
API = {
    "crn_config_reference": (C.c_int, [C.POINTER(Config)]),
    "crn_config_welch": (C.c_int, [C.POINTER(Config), C.c_int32, C.c_int32]),
    "crn_config_wideband": (C.c_int, [C.POINTER(Config), C.c_int32, C.c_int32, C.c_int32]),
    "crn_config_validate": (C.c_int, [C.POINTER(Config)]),
    "crn_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "crn_destroy": (C.c_int, [_P]),
    "crn_ring_acquire": (C.c_int, [_P, C.POINTER(_P)]),
    "crn_submit": (C.c_int, [_P, C.c_int32]),
    "crn_create_many": (C.c_int, [C.POINTER(Config), C.c_int32, C.POINTER(_P)]),
    "crn_submit_many": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32]),
    "crn_poll": (C.c_int, [_P, C.POINTER(Result)]),
    "crn_wait": (C.c_int, [_P, C.POINTER(Result)]),
    "crn_reset": (C.c_int, [_P]),
    "crn_sense_batch_host": (C.c_int, [_P, _P, C.c_int64, C.POINTER(Result)]),
    "crn_sense_batch_device": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "crn_synth_config_default": (C.c_int, [C.POINTER(SynthConfig), C.c_int32]),
    "crn_synth_generate_device": (C.c_int, [C.POINTER(SynthConfig), C.c_int32, _P, C.c_int64, C.c_int64, _P, _P]),
    "crn_synth_generate_streams_device": (C.c_int, [C.POINTER(SynthConfig), C.c_int32, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    "crn_fuse_masks_device": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "crn_ann_forward_device": (C.c_int, [C.POINTER(AnnWeights), C.c_double, _P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P]),
    "crn_ann_train_config_default": (C.c_int, [C.POINTER(AnnTrainConfig)]),
    "crn_ann_train_device": (C.c_int, [C.POINTER(AnnTrainConfig), _P, C.c_int32, _P, C.c_int64, C.POINTER(AnnWeights),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32, _P]),
    "crn_strerror": (C.c_char_p, [C.c_int]),
    "crn_last_error": (C.c_char_p, []),
    "crn_version": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crn_device_count": (C.c_int, []),
    "crn_launch_count": (C.c_int64, [_P]),
    "crn_get_kernel_info": (C.c_int, [_P, C.POINTER(KernelInfo)]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libcrnsense.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in API.items():
        try:
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        except AttributeError:
            if os.environ.get("CRN_LIB"):  # A/B against an older build variant: entry points it predates stay unbound
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class CrnError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib.crn_last_error().decode(errors="replace")
        super().__init__("%s: %s (%d)%s" % (where, lib.crn_strerror(status).decode(), status,
                                            ": " + msg if msg else ""))


def _check(status, where):
    if status != OK:
        raise CrnError(status, where)


def config_reference():
    """Reference-exact mode: N=512, K=10, rectangular, |X|, (sum)^2, 4 bands, the reference's MLP."""
    c = Config()
    _check(lib.crn_config_reference(C.byref(c)), "crn_config_reference")
    return c


def config_welch(nfft=1024, navg=64):
    """BASELINE config 2: Hann, |X|^2, K-frame Welch average, band plan scaled by nfft/512, MLP on."""
    c = Config()
    _check(lib.crn_config_welch(C.byref(c), nfft, navg), "crn_config_welch")
    return c


def config_wideband(nfft=8192, navg=64, nbands=64):
    """BASELINE config 3: nbands equal sub-channels, energy detection."""
    c = Config()
    _check(lib.crn_config_wideband(C.byref(c), nfft, navg, nbands), "crn_config_wideband")
    return c


def validate(cfg):
    return lib.crn_config_validate(C.byref(cfg))


def synth_config(group_samples, **kw):
    sc = SynthConfig()
    _check(lib.crn_synth_config_default(C.byref(sc), group_samples), "crn_synth_config_default")
    for k, v in kw.items():
        if k == "offsets_hz":
            for i in range(3):
                sc.offsets_hz[i] = v[i]
        else:
            setattr(sc, k, v)
    return sc


def device_count():
    return lib.crn_device_count()


def _ptr(t):
    """Device/host pointer of a torch tensor or numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class Sensor:
    """One sensing stream on one GPU (the GPU-side state of a CE_Predictive_Node instance)."""

    def __init__(self, cfg, device=None, _handle=None):
        self.cfg = cfg.copy()
        if device is not None:
            self.cfg.device = int(device)
        if _handle is not None:   # a member of a crn_create_many pool
            self._h = _handle
            return
        h = _P()
        _check(lib.crn_create(C.byref(self.cfg), C.byref(h)), "crn_create")
        self._h = h

    @classmethod
    def create_many(cls, cfg, n, device=None):
        """n streaming sensors of one configuration sharing one pinned ring (crn_create_many)."""
        c = cfg.copy()
        if device is not None:
            c.device = int(device)
        hs = (_P * n)()
        _check(lib.crn_create_many(C.byref(c), n, hs), "crn_create_many")
        return [cls(c, _handle=_P(hs[i])) for i in range(n)]

    @staticmethod
    def submit_many(sensors, nframes=1):
        """Commit nframes frames on every sensor; the members of one pool, in step, are sensed by ONE launch."""
        hs = (_P * len(sensors))(*[s._h for s in sensors])
        _check(lib.crn_submit_many(hs, len(sensors), nframes), "crn_submit_many")

    def ring_slot(self):
        """Address of the pinned slot the next frame goes to (crn_ring_acquire)."""
        slot = _P()
        _check(lib.crn_ring_acquire(self._h, C.byref(slot)), "crn_ring_acquire")
        return slot.value

    def close(self):
        if getattr(self, "_h", None):
            lib.crn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- streaming: one frame per USRP_RX_SAMPS event (CE_Predictive_Node.cpp:146-154) -------------
    def push_frame(self, frame):
        """Copy one frame of L complex64 samples into the pinned ring and commit it."""
        L = self.cfg.frame_len
        if self.cfg.iq_format == IQ_SC16:
            fr = np.ascontiguousarray(frame, dtype=np.int16).reshape(-1)
            if fr.shape != (2 * L,):
                raise ValueError("frame must hold exactly frame_len=%d (I,Q) int16 pairs" % L)
        else:
            fr = np.ascontiguousarray(frame, dtype=np.complex64)
            if fr.shape != (L,):
                raise ValueError("frame must hold exactly frame_len=%d complex samples" % L)
        slot = _P()
        _check(lib.crn_ring_acquire(self._h, C.byref(slot)), "crn_ring_acquire")
        C.memmove(slot, fr.ctypes.data, fr.nbytes)
        _check(lib.crn_submit(self._h, 1), "crn_submit")

    def poll(self):
        r = Result()
        st = lib.crn_poll(self._h, C.byref(r))
        if st == ERR_NOT_READY:
            return None
        _check(st, "crn_poll")
        return r

    def wait(self):
        r = Result()
        _check(lib.crn_wait(self._h, C.byref(r)), "crn_wait")
        return r

    def reset(self):
        _check(lib.crn_reset(self._h), "crn_reset")

    # -- batch ---------------------------------------------------------------------------------------
    def sense_host(self, iq, ngroups=None):
        """iq: host complex64 array (numpy, or a pinned torch tensor viewed as float32/complex64) holding
        ngroups*K frames.  Returns (feat[ng,nbands] f32, ann[ng,3] f64, decision[ng] i32, mask[ng] u64)."""
        gs = self.cfg.group_samples
        if hasattr(iq, "data_ptr"):
            nsamp = iq.numel() // (1 if iq.is_complex() else 2)
            ptr = C.c_void_p(iq.data_ptr())
        elif self.cfg.iq_format == IQ_SC16:
            iq = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
            nsamp = iq.size // 2
            ptr = C.c_void_p(iq.ctypes.data)
        else:
            iq = np.ascontiguousarray(iq, dtype=np.complex64)
            nsamp = iq.size
            ptr = C.c_void_p(iq.ctypes.data)
        if ngroups is None:
            ngroups = nsamp // gs
        if ngroups * gs > nsamp:
            raise ValueError("buffer holds %d samples, %d groups need %d" % (nsamp, ngroups, ngroups * gs))
        res = (Result * ngroups)()
        _check(lib.crn_sense_batch_host(self._h, ptr, ngroups, res), "crn_sense_batch_host")
        return results_to_arrays(res, self.cfg.nbands)

    def sense_host_raw(self, ptr, ngroups, res):
        """Timed e2e path: caller owns the (pinned) buffer and the Result array."""
        _check(lib.crn_sense_batch_host(self._h, ptr, ngroups, res), "crn_sense_batch_host")

    def sense_device(self, d_iq, ngroups, d_feat, d_ann=None, d_decision=None, d_mask=None, stream=0):
        """Asynchronous launch on `stream` (a raw cudaStream_t integer); all arguments are device tensors."""
        _check(lib.crn_sense_batch_device(self._h, _ptr(d_iq), ngroups, _ptr(d_feat), _ptr(d_ann),
                                          _ptr(d_decision), _ptr(d_mask), C.c_void_p(stream)),
               "crn_sense_batch_device")

    @property
    def launches(self):
        return lib.crn_launch_count(self._h)

    def kernel_info(self):
        ki = KernelInfo()
        _check(lib.crn_get_kernel_info(self._h, C.byref(ki)), "crn_get_kernel_info")
        return {f: (getattr(ki, f).decode() if f == "name" else getattr(ki, f)) for f, _ in KernelInfo._fields_}


def results_to_arrays(res, nbands):
    n = len(res)
    buf = np.frombuffer(res, dtype=np.dtype({
        "names": ["first_frame", "decision", "nfeat", "mask", "ann", "feat"],
        "formats": ["<u8", "<i4", "<i4", "<u8", ("<f8", 3), ("<f4", MAX_BANDS)],
        "offsets": [Result.first_frame.offset, Result.decision.offset, Result.nfeat.offset,
                    Result.occupancy_mask.offset, Result.ann_out.offset, Result.feat.offset],
        "itemsize": C.sizeof(Result)}), count=n)
    return (buf["feat"][:, :nbands].copy(), buf["ann"].copy(), buf["decision"].copy(), buf["mask"].copy())


def synth_generate(sc, d_iq, first_sample, nsamples, d_state=None, device=0, stream=0):
    """Fill a device tensor with the synthetic Markov-PU OFDM + AWGN capture (see crn_synth.cu)."""
    _check(lib.crn_synth_generate_device(C.byref(sc), device, _ptr(d_iq), first_sample, nsamples,
                                         _ptr(d_state), C.c_void_p(stream)), "crn_synth_generate_device")


def synth_generate_streams(sc, d_iq, first_stream, nstreams, samples_per_stream, d_state=None, device=0, stream=0):
    """Independent streams (multi-radio): stream first_stream + i -> d_iq[i * samples_per_stream : ...]."""
    _check(lib.crn_synth_generate_streams_device(C.byref(sc), device, _ptr(d_iq), first_stream, nstreams,
                                                 samples_per_stream, _ptr(d_state), C.c_void_p(stream)),
           "crn_synth_generate_streams_device")


def fuse_masks(d_masks, nradios, nslots, nbands, mode, d_fused, device=0, stream=0):
    """Cooperative hard-decision fusion of occupancy masks [nradios][nslots] -> [nslots] on the GPU."""
    _check(lib.crn_fuse_masks_device(_ptr(d_masks), nradios, nslots, nbands, mode, _ptr(d_fused), device,
                                     C.c_void_p(stream)), "crn_fuse_masks_device")


def shard_groups(ngroups, world_size, rank):
    """Contiguous block partition of decision groups over ranks (groups are independent, no
    collective on the data path).  Returns (first_group, count)."""
    base, rem = divmod(ngroups, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def ann_forward(weights, d_feat, n, feat_stride, d_out=None, d_decision=None, threshold=0.8, device=0, stream=0):
    """Batched MLP forward pass + first-match chain (CE_Predictive_Node.cpp:214-261) over device feature rows."""
    _check(lib.crn_ann_forward_device(C.byref(weights), threshold, _ptr(d_feat), n, feat_stride, _ptr(d_out),
                                      _ptr(d_decision), device, C.c_void_p(stream)), "crn_ann_forward_device")


def ann_train_config(**kw):
    tc = AnnTrainConfig()
    _check(lib.crn_ann_train_config_default(C.byref(tc)), "crn_ann_train_config_default")
    for k, v in kw.items():
        if k == "input_scale":
            for i in range(4):
                tc.input_scale[i] = v[i]
        else:
            setattr(tc, k, v)
    return tc


def ann_train(tc, d_feat, feat_stride, d_labels, n, weights=None, device=0, stream=0):
    """Batch back-propagation on the GPU from labelled feature rows.  Returns (weights, final_error, epochs_run);
    `weights` (AnnWeights) is the starting point when tc.init_range == 0."""
    w = AnnWeights()
    if weights is not None:
        C.memmove(C.byref(w), C.byref(weights), C.sizeof(AnnWeights))
    err, ep = C.c_double(0.0), C.c_int32(0)
    _check(lib.crn_ann_train_device(C.byref(tc), _ptr(d_feat), feat_stride, _ptr(d_labels), n, C.byref(w),
                                    C.byref(err), C.byref(ep), device, C.c_void_p(stream)), "crn_ann_train_device")
    return w, err.value, ep.value
