// Synthetic primary-user capture generated on the GPU: stands in for the USRP so
// the BASELINE workloads (1e9 samples = 8 GB) never cross PCIe.
//
// What it restates from the reference (test INPUT, not the hot path):
//   - PU waveform: ECR transmit path  src/extensible_cognitive_radio.cpp:883-949  (ofdmflexframegen with
//     the CRTS defaults: 64 subcarriers, cyclic prefix 16, taper 4 - src/crts.cpp:501-514; liquid's
//     default subcarrier allocation; unit power; soft gain -12 dB - ecr.cpp:59,892), generated at
//     1.4 MS/s (scenarios/predictive_model.cfg:39) and seen by the 13 MS/s receiver (:76).  Here the
//     OFDM symbol is evaluated directly as a continuous-time sum of its 50 used subcarriers at the
//     receiver's sample instants (an ideal 65/7 resampler).
//   - hopping: CE_PU_MARKOV_Chain_Tx.cpp:88-128 / CE_Random_Behaviour_PU.cpp:47-49, one draw per
//     dwell; the chain itself is sequential and tiny, so it is walked on the host and uploaded.
//   - complex AWGN at the stated in-band SNR (counter-hash Box-Muller).
//   - optionally an interferer node (src/interferer.cpp): the waveforms that need no modem - CW (:128-134),
//     uniform noise (:136-142), Gaussian noise exactly as coded, mean 5 / sigma 5 (:24,144-154) - generated at
//     the interferer's own rate, held to the receiver rate, scaled by its soft gain (:32,189), mixed to its
//     offset and gated by its duty cycle (:395-409).
//   - and the three that do (GMSK :156-221, root-raised-cosine QPSK :223-253, OFDM bursts :255-288), restated at the
//     level the sensing path sees (include/crnsense.h, enum crn_interferer); with pu_framed the PU's symbols follow
//     the flex-frame structure of transmit_frame (S0, S0, S1, header, payload; ecr.cpp:883-949).
// The CPU statement of the same definition is oracle/crn_oracle.c:crn_oracle_synth (test infrastructure).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "crn_internal.h"

namespace {

constexpr int SY_M = 64, SY_CP = 16, SY_TAPER = 4, SY_SYM = SY_M + SY_CP, SY_HALF = 25;
constexpr int SY_RATE_NUM = 7, SY_RATE_DEN = 65;  // 1.4e6 / 13e6

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__host__ __device__ inline unsigned long long stream_seed(unsigned long long seed, long long stream) {
  return mix64(seed ^ mix64((unsigned long long)stream * 0xD1B54A32D192ED03ull + 1));
}

int pu_next(int mode, int cur, int r) {
  if (mode == 2) return r % 3;                 // CE_Random_Behaviour_PU.cpp:47-49
  if (mode == 1) return r == 0 ? 0 : 1;        // as coded: CE_PU_MARKOV_Chain_Tx.cpp:104,114,123 are always true
  if (r == 0) return 0;                        // as documented: README.md:70-74
  if (cur == 0) return r < 4 ? 1 : 2;
  if (cur == 1) return r < 6 ? 1 : 2;
  return r < 3 ? 1 : 2;
}

// framed waveforms
constexpr int PU_FRAME_SYMS = 32;                       // S0 S0 S1 + 7 header + 22 payload
constexpr int OF_TAPER = 6, OF_FRAME_SYMS = 22;         // interferer.cpp:24; S0 S0 S1 + 7 header + 12 payload
constexpr int RRC_HLEN = 129, RRC_FRAME = 200;          // 2*2*32 + 1 taps (:61); 100 symbols x 2 samples (interferer.hpp:18-19)
constexpr int GM_SYMS = 1024, GM_SPS = 4, GM_PAD = 12;  // symbols per frame, samples per symbol after the x2 interpolator, :212
constexpr int GM_FRAME = GM_SYMS * GM_SPS + GM_PAD;

struct SynthParams {
  float2 *iq;
  const signed char *states;  // [nstreams][ndwell]
  int *group_state;           // [nstreams][groups per stream] or nullptr
  unsigned long long seed;    // base seed; stream i uses stream_seed(seed, first_stream + i)
  long long first_stream, sps, ndwell;  // samples per stream, dwells per stream
  long long first, n;         // first sample index inside a stream (single-stream mode), total samples
  long long dwell_samples;
  int group_samples;
  float gain, sigc;
  double cyc_per_sample[3];
  // interferer
  int intf_type;
  long long intf_period, intf_on;  // duty cycle in samples (period <= 0: always on)
  double intf_rate_ratio;          // interferer samples per receiver sample
  double intf_cyc_per_sample;
  float intf_gain;
  int pu_framed;
  float rrc_h[RRC_HLEN / 2 + 1];  // root raised cosine taps 0..64 (symmetric: h[i] = h[128 - i])
  float gm_q[20];                 // GMSK phase pulse q(x), x = -d + sub/4 - 1/2, index (d + 2) * 4 + sub
};

__device__ __forceinline__ void subcarrier(unsigned long long sseed, long long m, int k, float &re, float &im) {
  if (m < 0) { re = 0.f; im = 0.f; return; }
  const unsigned long long h = mix64(sseed ^ mix64((unsigned long long)m * 128ull + (unsigned long long)(k + 64)));
  const float g = 0.14142135623730950f;  // 1/sqrt(50)
  const int ak = k < 0 ? -k : k;
  if (((ak + 4) % 8) == 0) {
    re = (h & 1) ? g : -g;
    im = 0.f;
  } else {
    const float a = g * 0.70710678118654752f;
    re = (h & 1) ? a : -a;
    im = (h & 2) ? a : -a;
  }
}

// One subcarrier of symbol `fm` of a flex frame (liquid's ofdmflexframegen, recalled): S0 (fm 0, 1) carries a fixed
// +-1 sequence on the even used subcarriers only, S1 (fm 2) a fixed sequence on all of them, the header symbols BPSK,
// the payload symbols what `subcarrier` produces.  `uid` numbers the symbol for the payload/header hashes.
__device__ __forceinline__ void framed_subcarrier(unsigned long long sseed, long long uid, int fm, int nhdr, int k, float &re,
                                                  float &im) {
  if (uid < 0) { re = 0.f; im = 0.f; return; }
  const int ak = k < 0 ? -k : k;
  im = 0.f;
  if (fm < 2) {
    const unsigned long long h = mix64(0x5330ull * 0x9E3779B97F4A7C15ull + (unsigned long long)(k + 64));
    re = (ak & 1) ? 0.f : ((h & 1) ? 0.20412414523193151f : -0.20412414523193151f);  // 1/sqrt(24 even subcarriers)
  } else if (fm == 2) {
    const unsigned long long h = mix64(0x5331ull * 0x9E3779B97F4A7C15ull + (unsigned long long)(k + 64));
    re = (h & 1) ? 0.14142135623730950f : -0.14142135623730950f;
  } else if (fm < 3 + nhdr) {
    const unsigned long long h = mix64(sseed ^ mix64((unsigned long long)uid * 128ull + (unsigned long long)(k + 64)));
    re = (h & 1) ? 0.14142135623730950f : -0.14142135623730950f;
  } else {
    subcarrier(sseed, uid, k, re, im);
  }
}

// Sum of the used subcarriers of one OFDM symbol at phase z = exp(j 2 pi (tau - cp) / M), blended over the taper
// with the previous symbol's cyclic postfix.  FRAMED: symbol `m` sits at position m % frame_syms of its frame.
template <bool FRAMED>
__device__ __forceinline__ void ofdm_symbol_sum(unsigned long long sseed, long long m, int frame_syms, float zr, float zi,
                                                bool in_taper, float ramp, float &vr, float &vi) {
  float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f, pr = zr, pi = zi;
  const int fm = FRAMED ? (int)(m % frame_syms) : 0, fp = FRAMED ? (int)((m - 1 + frame_syms) % frame_syms) : 0;
#pragma unroll 5
  for (int k = 1; k <= SY_HALF; k++) {
    float xr, xi, yr, yi;
    if (FRAMED) {
      framed_subcarrier(sseed, m, fm, 7, k, xr, xi);
      framed_subcarrier(sseed, m, fm, 7, -k, yr, yi);
    } else {
      subcarrier(sseed, m, k, xr, xi);
      subcarrier(sseed, m, -k, yr, yi);
    }
    ar += xr * pr - xi * pi + yr * pr + yi * pi;
    ai += xr * pi + xi * pr - yr * pi + yi * pr;
    if (in_taper) {
      if (FRAMED) {
        framed_subcarrier(sseed, m - 1, fp, 7, k, xr, xi);
        framed_subcarrier(sseed, m - 1, fp, 7, -k, yr, yi);
      } else {
        subcarrier(sseed, m - 1, k, xr, xi);
        subcarrier(sseed, m - 1, -k, yr, yi);
      }
      br += xr * pr - xi * pi + yr * pr + yi * pi;
      bi += xr * pi + xi * pr - yr * pi + yi * pr;
    }
    const float nr = pr * zr - pi * zi;
    pi = pr * zi + pi * zr;
    pr = nr;
  }
  vr = ramp * ar + (1.0f - ramp) * br;
  vi = ramp * ai + (1.0f - ramp) * bi;
}

// Interferer sample `im` (at the interferer's own rate) of the framed waveforms.
__device__ __forceinline__ void intf_modem_sample(const SynthParams &p, unsigned long long sseed, unsigned long long im,
                                                  float &re, float &imag) {
  re = 0.f;
  imag = 0.f;
  if (p.intf_type == CRN_INTF_RRC) {
    // interferer.cpp:223-253: y[j] = sum_i h[i] x[j - i], x nonzero at the even samples of the frame
    const unsigned long long F = im / RRC_FRAME;
    const int j = (int)(im % RRC_FRAME);
    const int nlo = j > RRC_HLEN - 1 ? (j - (RRC_HLEN - 1) + 1) / 2 : 0, nhi = j / 2;
    for (int n = nlo; n <= nhi; n++) {
      const int tap = j - 2 * n;
      const float h = p.rrc_h[tap <= RRC_HLEN / 2 ? tap : RRC_HLEN - 1 - tap];
      const unsigned long long hs = mix64(sseed ^ mix64(0x5252430000000000ull + F * 128ull + (unsigned long long)n));
      re = fmaf((hs & 1) ? 0.25f : -0.25f, h, re);
      imag = fmaf((hs & 2) ? 0.25f : -0.25f, h, imag);
    }
  } else if (p.intf_type == CRN_INTF_GMSK) {
    // interferer.cpp:156-221: constant envelope, phase = pi/2 * sum_n b_n q(t - n - 1/2)
    const unsigned long long F = im / GM_FRAME;
    const int j = (int)(im % GM_FRAME);
    if (j >= GM_SYMS * GM_SPS) return;  // padding between frames (:212-219)
    const int n0 = j / GM_SPS, sub = j % GM_SPS;
    auto word = [&](int w) -> unsigned {
      return (unsigned)mix64(sseed ^ mix64(0x474D534B00000000ull + F * 64ull + (unsigned long long)w));
    };
    // bits up to n0 - 3 have turned the phase by a full +-pi/2 each: count them a word at a time
    int turns = 0;
    const int full = n0 - 3;  // last fully integrated bit
    if (full >= 0) {
      const int wl = full >> 5;
      for (int w = 0; w < wl; w++) turns += 2 * __popc(word(w)) - 32;
      const int nb = (full & 31) + 1;
      const unsigned msk = nb == 32 ? 0xFFFFFFFFu : ((1u << nb) - 1u);
      turns += 2 * __popc(word(wl) & msk) - nb;
    }
    float frac = 0.f;
#pragma unroll
    for (int d = -2; d <= 2; d++) {
      const int n = n0 + d;
      if (n >= 0 && n < GM_SYMS) {
        const float b = ((word(n >> 5) >> (n & 31)) & 1u) ? 1.0f : -1.0f;
        frac = fmaf(b, p.gm_q[(d + 2) * 4 + sub], frac);
      }
    }
    const float quarter = (float)(turns & 3) + frac;  // phase in quarter turns
    sincosf(1.5707963267948966f * quarter, &imag, &re);
  } else if (p.intf_type == CRN_INTF_OFDM) {
    // interferer.cpp:255-288: symbols of a 64-subcarrier flex frame at the interferer's rate, integer sample phases
    const long long m = (long long)(im / SY_SYM);
    const int tau = (int)(im % SY_SYM);
    float zr, zi;
    sincosf(6.283185307179586f * (float)(tau - SY_CP) * (1.0f / SY_M), &zi, &zr);
    float ramp = 1.0f;
    const bool in_taper = tau < OF_TAPER;
    if (in_taper) {
      const float sn = sinf(1.5707963267948966f * ((float)tau + 0.5f) * (1.0f / OF_TAPER));
      ramp = sn * sn;
    }
    // frame structure: header 7 symbols, 12 payload symbols; payload/header hashes are keyed by the global symbol index
    float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f, pr = zr, pi = zi;
    const int fm = (int)(m % OF_FRAME_SYMS), fp = (int)((m - 1 + OF_FRAME_SYMS) % OF_FRAME_SYMS);
    const unsigned long long iseed = sseed ^ 0x4F46444D4F46444Dull;
    for (int k = 1; k <= SY_HALF; k++) {
      float xr, xi, yr, yi;
      framed_subcarrier(iseed, m, fm, 7, k, xr, xi);
      framed_subcarrier(iseed, m, fm, 7, -k, yr, yi);
      ar += xr * pr - xi * pi + yr * pr + yi * pi;
      ai += xr * pi + xi * pr - yr * pi + yi * pr;
      if (in_taper) {
        framed_subcarrier(iseed, m - 1, fp, 7, k, xr, xi);
        framed_subcarrier(iseed, m - 1, fp, 7, -k, yr, yi);
        br += xr * pr - xi * pi + yr * pr + yi * pi;
        bi += xr * pi + xi * pr - yr * pi + yi * pr;
      }
      const float nr = pr * zr - pi * zi;
      pi = pr * zi + pi * zr;
      pr = nr;
    }
    re = ramp * ar + (1.0f - ramp) * br;
    imag = ramp * ai + (1.0f - ramp) * bi;
  }
}

__global__ void __launch_bounds__(256) synth_kernel(const SynthParams p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const long long si = i / p.sps;                 // stream slot in this call
    const long long s = p.first + (i - si * p.sps);  // sample index inside the stream
    const unsigned long long sseed = stream_seed(p.seed, p.first_stream + si);
    const int ch = p.states[si * p.ndwell + s / p.dwell_samples];
    if (p.group_state && (s % p.group_samples) == 0)
      p.group_state[si * (p.sps / p.group_samples) + s / p.group_samples - p.first / p.group_samples] = ch;
    const long long un = s * SY_RATE_NUM;
    const long long m = un / ((long long)SY_RATE_DEN * SY_SYM);
    const float tau = (float)(un % ((long long)SY_RATE_DEN * SY_SYM)) / (float)SY_RATE_DEN;
    const float th = (tau - (float)SY_CP) * (1.0f / SY_M);
    float zr, zi;
    sincosf(6.283185307179586f * th, &zi, &zr);
    float ramp = 1.0f;
    const bool in_taper = tau < (float)SY_TAPER;
    if (in_taper) {
      const float sn = sinf(1.5707963267948966f * tau * (1.0f / SY_TAPER));
      ramp = sn * sn;
    }
    float sumr, sumi;
    if (p.pu_framed) ofdm_symbol_sum<true>(sseed, m, PU_FRAME_SYMS, zr, zi, in_taper, ramp, sumr, sumi);
    else ofdm_symbol_sum<false>(sseed, m, 1, zr, zi, in_taper, ramp, sumr, sumi);
    const float vr = p.gain * sumr, vi = p.gain * sumi;
    const double cyc = (double)s * p.cyc_per_sample[ch];
    const float ph = (float)(cyc - floor(cyc));
    float cr, ci;
    sincosf(6.283185307179586f * ph, &ci, &cr);
    float outr = vr * cr - vi * ci, outi = vr * ci + vi * cr;
    const unsigned long long h = mix64(sseed ^ mix64(2ull * (unsigned long long)s + 1ull));
    const float u1 = (float)((h >> 40) + 1ull) * (1.0f / 16777216.0f);
    const float u2 = (float)((h >> 16) & 0xFFFFFFull) * (1.0f / 16777216.0f);
    const float rad = p.sigc * sqrtf(-2.0f * logf(u1));
    float nc, ns;
    sincosf(6.283185307179586f * u2, &ns, &nc);
    outr += rad * nc;
    outi += rad * ns;
    if (p.intf_type != CRN_INTF_NONE && (p.intf_period <= 0 || (s % p.intf_period) < p.intf_on)) {
      const unsigned long long im = (unsigned long long)((double)s * p.intf_rate_ratio);  // its sample index
      float br2 = 0.5f, bi2 = 0.5f;  // CW
      if (p.intf_type >= CRN_INTF_GMSK) {
        intf_modem_sample(p, sseed, im, br2, bi2);
      } else if (p.intf_type != CRN_INTF_CW) {
        const unsigned long long hi = mix64(sseed ^ mix64(0x1F7E2A5C00000000ull + 2ull * im));
        const float v1 = (float)(hi >> 40) * (1.0f / 16777216.0f), v2 = (float)((hi >> 16) & 0xFFFFFFull) * (1.0f / 16777216.0f);
        if (p.intf_type == CRN_INTF_NOISE) {
          br2 = 0.5f * v1 - 0.25f;
          bi2 = 0.5f * v2 - 0.25f;
        } else {  // independent N(5, 5) draws for the two components: the two Box-Muller outputs
          const float r5 = 5.0f * sqrtf(-2.0f * logf((float)((hi >> 40) + 1ull) * (1.0f / 16777216.0f)));
          float gs, gc;
          sincosf(6.283185307179586f * v2, &gs, &gc);
          br2 = 5.0f + r5 * gc;
          bi2 = 5.0f + r5 * gs;
        }
      }
      const double icyc = (double)s * p.intf_cyc_per_sample;
      float jr, ji;
      sincosf(6.283185307179586f * (float)(icyc - floor(icyc)), &ji, &jr);
      outr += p.intf_gain * (br2 * jr - bi2 * ji);
      outi += p.intf_gain * (br2 * ji + bi2 * jr);
    }
    p.iq[i] = make_float2(outr, outi);
  }
}

}  // namespace

namespace {
int synth_launch(const crn_synth_config *sc, int32_t device, void *d_iq, int64_t first_stream, int64_t nstreams,
                 int64_t first_sample, int64_t sps, int32_t *d_state, void *cuda_stream) {
  if (sps == 0 || nstreams == 0) return CRN_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cudaStream_t st = (cudaStream_t)cuda_stream;

  SynthParams p;
  memset(&p, 0, sizeof(p));
  p.iq = (float2 *)d_iq;
  p.group_state = d_state;
  p.seed = sc->seed;
  p.first_stream = first_stream;
  p.sps = sps;
  p.first = first_sample;
  p.n = sps * nstreams;
  p.dwell_samples = (long long)sc->dwell_groups * sc->group_samples;
  p.group_samples = sc->group_samples;
  p.gain = (float)pow(10.0, sc->pu_gain_db / 20.0);
  const double ps = pow(10.0, sc->pu_gain_db / 10.0);
  const double bocc = (2.0 * SY_HALF + 1.0) / SY_M * sc->pu_rate;
  const double sigma2 = ps * sc->fs / (bocc * pow(10.0, sc->snr_db / 10.0));
  p.sigc = (float)sqrt(sigma2 / 2.0);
  for (int c = 0; c < 3; c++) p.cyc_per_sample[c] = sc->offsets_hz[c] / sc->fs;
  p.intf_type = sc->intf_type;
  p.intf_period = (long long)sc->intf_period_groups * sc->group_samples;
  p.intf_on = (long long)llround(sc->intf_duty * (double)p.intf_period);
  p.intf_rate_ratio = sc->intf_rate / sc->fs;
  p.intf_cyc_per_sample = sc->intf_offset_hz / sc->fs;
  p.intf_gain = (float)pow(10.0, sc->intf_gain_db / 20.0);
  p.pu_framed = sc->pu_framed;
  {  // liquid_firdes_rrcos(k = 2, m = 32, beta = 0.35, dt = 0) as recalled: taps 0..64 (the rest by symmetry)
    const double beta = 0.35;
    for (int i = 0; i <= RRC_HLEN / 2; i++) {
      const double z = (double)i / 2.0 - 32.0, g = 1.0 - 16.0 * beta * beta * z * z;
      double h;
      if (fabs(z) < 1e-5) h = 1.0 - beta + 4.0 * beta / M_PI;
      else if (fabs(g) < 1e-5) h = beta / sqrt(2.0) * ((1.0 + 2.0 / M_PI) * sin(M_PI / (4.0 * beta)) + (1.0 - 2.0 / M_PI) * cos(M_PI / (4.0 * beta)));
      else h = (4.0 * beta / (M_PI * g)) * (cos((1.0 + beta) * M_PI * z) + sin((1.0 - beta) * M_PI * z) / (4.0 * beta * z));
      p.rrc_h[i] = (float)h;
    }
    // GMSK phase pulse, BT = 0.5: q(x) = I(x + 1/2) - I(x - 1/2), I(u) = u Phi(c u) + phi(c u) / c, c = 2 pi BT / sqrt(ln 2)
    const double c = 2.0 * M_PI * 0.5 / sqrt(log(2.0));
    auto I = [&](double u) { return u * 0.5 * erfc(-c * u / sqrt(2.0)) + exp(-0.5 * c * c * u * u) / (sqrt(2.0 * M_PI) * c); };
    for (int d = -2; d <= 2; d++)
      for (int sub = 0; sub < 4; sub++) {
        const double x = -(double)d + (double)sub / 4.0 - 0.5;
        p.gm_q[(d + 2) * 4 + sub] = (float)(I(x + 0.5) - I(x - 0.5));
      }
  }

  // walk every stream's hop chain on the host from dwell 0 (sequential by definition) and upload it
  const long long ndwell = (first_sample + sps + p.dwell_samples - 1) / p.dwell_samples;
  p.ndwell = ndwell;
  std::vector<signed char> states((size_t)(ndwell * nstreams));
  for (long long si = 0; si < nstreams; si++) {
    const unsigned long long sseed = stream_seed(sc->seed, first_stream + si);
    int cur = 0;  // dwell 0 on CH1: tx_freq = 833e6, scenarios/predictive_model.cfg:37
    for (long long d = 0; d < ndwell; d++) {
      if (d > 0) {
        const unsigned long long h = mix64(sseed ^ (0xA5A5A5A5ull + (unsigned long long)d * 0x2545F4914F6CDD1Dull));
        const int r = (sc->hop_mode == 2) ? (int)((h >> 33) % 3) : (int)((h >> 33) % 10);
        cur = pu_next(sc->hop_mode, cur, r);
      }
      states[(size_t)(si * ndwell + d)] = (signed char)cur;
    }
  }
  signed char *d_states = nullptr;
  e = cudaMallocAsync(&d_states, states.size(), st);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "cudaMallocAsync: %s", cudaGetErrorString(e));
  // pageable source: the copy is staged before the call returns, `states` may go out of scope
  e = cudaMemcpyAsync(d_states, states.data(), states.size(), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
  p.states = d_states;

  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device);
  long long blocks = (p.n + 255) / 256;
  const long long cap = (long long)dev_sms * 8;
  if (blocks > cap) blocks = cap;
  synth_kernel<<<(int)blocks, 256, 0, st>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "synth kernel launch: %s", cudaGetErrorString(e));
  e = cudaFreeAsync(d_states, st);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "cudaFreeAsync: %s", cudaGetErrorString(e));
  return CRN_OK;
}
}  // namespace

extern "C" int crn_synth_generate_device(const crn_synth_config *sc, int32_t device, void *d_iq,
                                         int64_t first_sample, int64_t nsamples, int32_t *d_state,
                                         void *cuda_stream) {
  if (!sc || !d_iq || first_sample < 0 || nsamples < 0 || sc->group_samples < 1 || sc->dwell_groups < 1)
    return crn::fail(CRN_ERR_INVALID, "crn_synth_generate_device: bad argument");
  if (sc->intf_type < CRN_INTF_NONE || sc->intf_type > CRN_INTF_OFDM ||
      (sc->intf_type != CRN_INTF_NONE && (!(sc->intf_rate > 0.0) || sc->intf_duty < 0.0 || sc->intf_duty > 1.0)))
    return crn::fail(CRN_ERR_INVALID, "crn_synth_generate_device: bad interferer settings");
  if (d_state && ((first_sample % sc->group_samples) != 0 || (nsamples % sc->group_samples) != 0))
    return crn::fail(CRN_ERR_INVALID, "first_sample and nsamples must be group aligned when d_state is requested");
  return synth_launch(sc, device, d_iq, 0, 1, first_sample, nsamples, d_state, cuda_stream);
}

extern "C" int crn_synth_generate_streams_device(const crn_synth_config *sc, int32_t device, void *d_iq,
                                                 int64_t first_stream, int64_t nstreams,
                                                 int64_t samples_per_stream, int32_t *d_state,
                                                 void *cuda_stream) {
  if (!sc || !d_iq || first_stream < 0 || nstreams < 0 || samples_per_stream < 0 || sc->group_samples < 1 ||
      sc->dwell_groups < 1)
    return crn::fail(CRN_ERR_INVALID, "crn_synth_generate_streams_device: bad argument");
  if (sc->intf_type < CRN_INTF_NONE || sc->intf_type > CRN_INTF_OFDM ||
      (sc->intf_type != CRN_INTF_NONE && (!(sc->intf_rate > 0.0) || sc->intf_duty < 0.0 || sc->intf_duty > 1.0)))
    return crn::fail(CRN_ERR_INVALID, "crn_synth_generate_streams_device: bad interferer settings");
  if (d_state && (samples_per_stream % sc->group_samples) != 0)
    return crn::fail(CRN_ERR_INVALID, "samples_per_stream must be group aligned when d_state is requested");
  return synth_launch(sc, device, d_iq, first_stream, nstreams, 0, samples_per_stream, d_state, cuda_stream);
}
