// libcrnsense: configuration helpers (host only, no CUDA).
//
// These replace the reference engine's compile-time constants and literals:
//   fft_length / fft_averaging            CE_Predictive_Node.hpp:31-32
//   the hand-picked bin ranges            CE_Predictive_Node.cpp:173-191
//   the 25 + 18 MLP weights               CE_Predictive_Node.cpp:78-120
//   the 0.8 decision threshold            CE_Predictive_Node.cpp:245,250,255
#include <cstdio>
#include <cstring>

#include "crn_internal.h"

namespace {

// Bin plan of the reference at N = 512, as half-open ranges, in ANN-input order
// (Features_Buffer = {0, NF^2, CH1, CH2, CH3}, CE_Predictive_Node.cpp:200).
struct RefSeg { int band, lo, hi; };
const RefSeg kRefSegs[] = {
    {0, 300, 310},  // NF   .cpp:189-191
    {1, 0, 16},     // CH1  .cpp:173-175  (833 MHz, DC side)
    {1, 496, 511},  // CH1  .cpp:177-179  (bin 511 is NOT included upstream)
    {2, 55, 85},    // CH2  .cpp:181-183  (835 MHz)
    {3, 189, 222},  // CH3  .cpp:185-187  (838 MHz)
};

void load_reference_weights(crn_config *c) {
  memset(c->ann_wih, 0, sizeof(c->ann_wih));
  memset(c->ann_who, 0, sizeof(c->ann_who));
  // WeightIH[i][j], listed per hidden unit j: {bias(i=0), i=1..4}   (.cpp:78-102)
  static const double ih[5][5] = {
      {-0.188208, -0.106634, 0.005650, -0.057578, 0.092680},
      {-0.170684, -0.415470, 0.741944, 0.621154, 0.809336},
      {-0.024726, 0.309261, 0.006133, -0.048268, -0.010821},
      {0.001448, 0.159974, -0.620100, -0.249186, -0.546496},
      {0.015983, 0.212781, 0.669892, 0.734475, 0.609384},
  };
  // WeightHO[j][k], listed per output k: {bias(j=0), j=1..5}        (.cpp:103-120)
  static const double ho[3][6] = {
      {-7.033320, 10.857465, -6.848443, 17.053079, 0.087664, -6.552455},
      {2.726400, -18.452471, 2.053071, -13.375309, -0.269499, 2.655529},
      {-2.590206, 15.609466, -2.929559, -15.703407, 0.407028, -2.552555},
  };
  for (int j = 1; j <= CRN_ANN_HIDDEN; j++)
    for (int i = 0; i <= CRN_ANN_INPUTS; i++) c->ann_wih[i][j] = ih[j - 1][i];
  for (int k = 1; k <= CRN_ANN_OUTPUTS; k++)
    for (int j = 0; j <= CRN_ANN_HIDDEN; j++) c->ann_who[j][k] = ho[k - 1][j];
  c->ann_threshold = 0.8;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" {

int crn_config_reference(crn_config *c) {
  if (!c) return crn::fail(CRN_ERR_INVALID, "crn_config_reference: null config");
  memset(c, 0, sizeof(*c));
  c->nfft = 512;
  c->frame_len = 512;
  c->frame_stride = 0;
  c->navg = 10;
  c->window = CRN_WINDOW_RECT;
  c->detector = CRN_DET_MAG;
  c->postop = CRN_POST_SQUARE_OF_SUM;
  c->decide = CRN_DECIDE_ANN;
  c->nbands = 4;
  c->nsegs = (int)(sizeof(kRefSegs) / sizeof(kRefSegs[0]));
  for (int s = 0; s < c->nsegs; s++) {
    c->segs[s].band = kRefSegs[s].band;
    c->segs[s].lo = kRefSegs[s].lo;
    c->segs[s].hi = kRefSegs[s].hi;
  }
  load_reference_weights(c);
  c->energy_factor = 4.0;
  c->device = 0;
  c->ring_slots = 4;
  c->iq_format = CRN_IQ_CF32;
  return CRN_OK;
}

int crn_config_welch(crn_config *c, int32_t nfft, int32_t navg) {
  if (!c) return crn::fail(CRN_ERR_INVALID, "crn_config_welch: null config");
  if (!is_pow2(nfft) || nfft < 256 || nfft > 8192 || navg < 1)
    return crn::fail(CRN_ERR_INVALID, "crn_config_welch: nfft must be a power of two in [256,8192], navg >= 1");
  crn_config_reference(c);
  c->nfft = nfft;
  c->frame_len = nfft;
  c->navg = navg;
  c->window = CRN_WINDOW_HANN;
  c->detector = CRN_DET_MAGSQ;
  c->postop = CRN_POST_SUM;
  // same Hz edges: bin indices scale with nfft / 512 (rounded down at N = 256, where the bins are twice as wide)
  for (int s = 0; s < c->nsegs; s++) {
    c->segs[s].lo = (int)((long long)c->segs[s].lo * nfft / 512);
    c->segs[s].hi = (int)((long long)c->segs[s].hi * nfft / 512);
  }
  return CRN_OK;
}

int crn_config_wideband(crn_config *c, int32_t nfft, int32_t navg, int32_t nbands) {
  if (!c) return crn::fail(CRN_ERR_INVALID, "crn_config_wideband: null config");
  if (!is_pow2(nfft) || nfft < 256 || nfft > 8192 || navg < 1 || nbands < 1 ||
      nbands > CRN_MAX_BANDS || nfft % nbands != 0)
    return crn::fail(CRN_ERR_INVALID, "crn_config_wideband: bad nfft/navg/nbands");
  crn_config_reference(c);
  c->nfft = nfft;
  c->frame_len = nfft;
  c->navg = navg;
  c->window = CRN_WINDOW_HANN;
  c->detector = CRN_DET_MAGSQ;
  c->postop = CRN_POST_SUM;
  c->decide = CRN_DECIDE_ENERGY;
  c->nbands = nbands;
  c->nsegs = nbands;
  const int w = nfft / nbands;
  for (int b = 0; b < nbands; b++) {
    c->segs[b].band = b;
    c->segs[b].lo = b * w;
    c->segs[b].hi = (b + 1) * w;
  }
  return CRN_OK;
}

int crn_config_validate(const crn_config *c) {
  if (!c) return crn::fail(CRN_ERR_INVALID, "null config");
  if (!is_pow2(c->nfft)) return crn::fail(CRN_ERR_INVALID, "nfft must be a power of two");
  if (c->nfft < 256 || c->nfft > 8192)
    return crn::fail(CRN_ERR_UNSUPPORTED, "nfft must be in [256, 8192]");
  if (c->frame_len < 1 || c->frame_len > c->nfft)
    return crn::fail(CRN_ERR_INVALID,
                     "frame_len must be in [1, nfft] (the reference overruns buffer[512] when L > N; "
                     "this library refuses instead)");
  if (c->frame_stride != 0 && c->frame_stride < c->frame_len)
    return crn::fail(CRN_ERR_INVALID, "frame_stride must be 0 or >= frame_len");
  if (c->navg < 1 || c->navg > 65536) return crn::fail(CRN_ERR_INVALID, "navg must be in [1, 65536]");
  if (c->window != CRN_WINDOW_RECT && c->window != CRN_WINDOW_HANN)
    return crn::fail(CRN_ERR_INVALID, "unknown window");
  if (c->detector != CRN_DET_MAG && c->detector != CRN_DET_MAGSQ)
    return crn::fail(CRN_ERR_INVALID, "unknown detector");
  if (c->postop != CRN_POST_SQUARE_OF_SUM && c->postop != CRN_POST_SUM && c->postop != CRN_POST_SUM_DB)
    return crn::fail(CRN_ERR_INVALID, "unknown postop");
  if (c->decide < CRN_DECIDE_NONE || c->decide > CRN_DECIDE_ENERGY)
    return crn::fail(CRN_ERR_INVALID, "unknown decide mode");
  if (c->nbands < 1 || c->nbands > CRN_MAX_BANDS) return crn::fail(CRN_ERR_INVALID, "nbands out of range");
  if (c->nsegs < 1 || c->nsegs > CRN_MAX_SEGS) return crn::fail(CRN_ERR_INVALID, "nsegs out of range");
  if (c->decide == CRN_DECIDE_ANN && c->nbands < CRN_ANN_INPUTS)
    return crn::fail(CRN_ERR_INVALID, "ANN decision needs >= 4 bands (NF, CH1, CH2, CH3)");
  for (int s = 0; s < c->nsegs; s++) {
    const crn_seg &g = c->segs[s];
    if (g.band < 0 || g.band >= c->nbands || g.lo < 0 || g.hi > c->nfft || g.lo >= g.hi)
      return crn::fail(CRN_ERR_INVALID, "segment out of range");
  }
  if (c->ring_slots != 0 && c->ring_slots < 2) return crn::fail(CRN_ERR_INVALID, "ring_slots must be >= 2");
  if (c->iq_format != CRN_IQ_CF32 && c->iq_format != CRN_IQ_SC16) return crn::fail(CRN_ERR_INVALID, "unknown iq_format");
  return CRN_OK;
}

int crn_synth_config_default(crn_synth_config *sc, int32_t group_samples) {
  if (!sc || group_samples < 1) return crn::fail(CRN_ERR_INVALID, "crn_synth_config_default: bad argument");
  memset(sc, 0, sizeof(*sc));
  sc->seed = 12;  // echoes srand(12), src/crts_cognitive_radio.cpp:754
  sc->fs = 13e6;
  sc->pu_rate = 1.4e6;
  sc->offsets_hz[0] = 0.0;
  sc->offsets_hz[1] = 2e6;
  sc->offsets_hz[2] = 5e6;
  sc->snr_db = 10.0;
  sc->pu_gain_db = -12.0;
  sc->hop_mode = 0;
  sc->dwell_groups = 64;
  sc->group_samples = group_samples;
  sc->intf_type = CRN_INTF_NONE;
  sc->intf_period_groups = 0;
  sc->intf_offset_hz = 0.0;
  sc->intf_rate = 1e6;      // scenarios/interferer defaults to tx_rate 1e6 in src/crts.cpp
  sc->intf_gain_db = -3.0;  // interferer.cpp:32
  sc->intf_duty = 1.0;      // interferer.cpp:29
  return CRN_OK;
}

const char *crn_strerror(int status) {
  switch (status) {
    case CRN_OK: return "ok";
    case CRN_ERR_INVALID: return "invalid argument or configuration";
    case CRN_ERR_NO_DEVICE: return "no usable CUDA device";
    case CRN_ERR_CUDA: return "CUDA runtime error";
    case CRN_ERR_NOMEM: return "out of memory";
    case CRN_ERR_OVERRUN: return "ring overrun";
    case CRN_ERR_NOT_READY: return "no result ready";
    case CRN_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

const char *crn_last_error(void) { return crn::last_error(); }

int crn_version(int32_t *major, int32_t *minor) {
  if (major) *major = CRN_VERSION_MAJOR;
  if (minor) *minor = CRN_VERSION_MINOR;
  return CRN_OK;
}

}  // extern "C"

namespace crn {
namespace {
thread_local char g_err[512] = "";
}
int fail(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}
const char *last_error() { return g_err; }
}  // namespace crn
