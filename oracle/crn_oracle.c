/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into libcrnsense or any product path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Oracle "port": plain-C restatement of the reference's sensing algorithm
 *     cognitive_engines/CE_Predictive_Node/CE_Predictive_Node.cpp:146-261
 * parametrised by the same crn_config the GPU library takes, so the configurations the reference
 * cannot express (N != 512, Hann, |X|^2, K = 64, 64 sub-channels) are the same loop nest with the
 * options flipped.  In reference-exact mode (crn_config_reference) it is checked BIT-EXACT against
 * oracle O1 (the unmodified engine object driven by oracle/ref_harness.cpp) in tests/test_oracle.py.
 *
 * PARITY UNPINNED by the reference itself: it ships no tests, golden vectors or recorded IQ, and its FFT
 * (liquid-dsp a4d7c80d3) is an absent third-party dependency restated in oracle/liquid_fft_restated.c.
 * Pins that do come from the reference: the 43 weight literals (.cpp:78-120), the bin table
 * (.cpp:173-190), the order of operations, and the unmodified engine object itself.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/crnsense.h"

typedef float _Complex cf32;
typedef struct fftplan_s *fftplan;
fftplan fft_create_plan(unsigned int n, cf32 *x, cf32 *y, int dir, int flags);
void fft_destroy_plan(fftplan p);
void fft_execute(fftplan p);

/* liquid-dsp's hann(n, N) = 0.5 - 0.5 cos(2 pi n / (N-1)) (symmetric; src/math/src/windows.c,
   recalled - liquid is absent).  The reference engine itself applies no window. */
float crn_oracle_hann(int n, int N) {
  return 0.5f - 0.5f * cosf((float)(2.0 * M_PI * (double)n) / (float)(N - 1));
}

/* One decision group: frames[K][stride] -> feat[nbands], ann[3], decision, mask. */
static void sense_group(const crn_config *c, const float *iq, fftplan plan, cf32 *buf, cf32 *spec,
                        float *avg, const float *win, float *feat, double *ann, int32_t *decision,
                        uint64_t *mask) {
  const int N = c->nfft, L = c->frame_len, K = c->navg;
  const int stride = c->frame_stride > 0 ? c->frame_stride : L;
  memset(avg, 0, sizeof(float) * N); /* .cpp:39,287 */
  const int sc16 = (c->iq_format == CRN_IQ_SC16);
  for (int k = 0; k < K; k++) {
    const float *fr = iq + 2 * (size_t)k * stride;
    const int16_t *fr16 = (const int16_t *)iq + 2 * (size_t)k * stride;
    /* .cpp:149: memcpy of L samples into buffer[N]; the tail stays zero */
    for (int n = 0; n < N; n++) {
      if (n < L) {
        /* sc16 wire samples are what UHD converts to fc32 before the engine sees them: value = int16/32768 */
        float re = sc16 ? (float)fr16[2 * n] * (1.0f / 32768.0f) : fr[2 * n];
        float im = sc16 ? (float)fr16[2 * n + 1] * (1.0f / 32768.0f) : fr[2 * n + 1];
        if (win) { re *= win[n]; im *= win[n]; }
        buf[n] = re + im * _Complex_I;
      } else {
        buf[n] = 0;
      }
    }
    fft_execute(plan); /* .cpp:150 */
    for (int i = 0; i < N; i++) { /* .cpp:152-154 */
      float d;
      if (c->detector == CRN_DET_MAG) {
        d = cabsf(spec[i]);
      } else {
        const float re = crealf(spec[i]), im = cimagf(spec[i]);
        d = re * re + im * im;
      }
      avg[i] += d / (float)K;
    }
  }
  /* .cpp:163-197: per-band sums in the listed order, then the power feature */
  float m[CRN_MAX_BANDS];
  for (int b = 0; b < c->nbands; b++) m[b] = 0.0f;
  for (int s = 0; s < c->nsegs; s++)
    for (int i = c->segs[s].lo; i < c->segs[s].hi; i++) m[c->segs[s].band] += avg[i];
  float lin[CRN_MAX_BANDS]; /* linear features: what the MLP / energy detector see */
  for (int b = 0; b < c->nbands; b++) {
    lin[b] = (c->postop == CRN_POST_SQUARE_OF_SUM) ? m[b] * m[b] : m[b];
    feat[b] = (c->postop == CRN_POST_SUM_DB) ? 10.0f * log10f(lin[b]) : lin[b];
  }

  double out[CRN_ANN_OUTPUTS + 1] = {0, 0, 0, 0};
  int dec = 0;
  uint64_t msk = 0;
  if (c->decide == CRN_DECIDE_ANN) {
    /* .cpp:200: Features_Buffer = {0, NF^2, CH1, CH2, CH3} == features 0..3 */
    double F[CRN_ANN_INPUTS + 1] = {0, lin[0], lin[1], lin[2], lin[3]};
    double H[CRN_ANN_HIDDEN + 1];
    for (int j = 1; j <= CRN_ANN_HIDDEN; j++) { /* .cpp:214-220 */
      double sum = c->ann_wih[0][j];
      for (int i = 1; i <= CRN_ANN_INPUTS; i++) sum += F[i] * c->ann_wih[i][j];
      H[j] = 1.0 / (1.0 + exp(-sum));
    }
    for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) { /* .cpp:229-235 */
      double sum = c->ann_who[0][k];
      for (int j = 1; j <= CRN_ANN_HIDDEN; j++) sum += H[j] * c->ann_who[j][k];
      out[k] = 1.0 / (1.0 + exp(-sum));
    }
    /* .cpp:245-261 first-match chain */
    if (out[1] >= c->ann_threshold) dec = CRN_CH1_OCCUPIED;
    else if (out[2] >= c->ann_threshold) dec = CRN_CH2_OCCUPIED;
    else if (out[3] >= c->ann_threshold) dec = CRN_CH3_OCCUPIED;
    else dec = CRN_ALL_BUSY;
    msk = dec ? (1ull << (dec - 1)) : 0;
  } else if (c->decide == CRN_DECIDE_ENERGY) {
    float mn = lin[0];
    for (int b = 1; b < c->nbands; b++) mn = lin[b] < mn ? lin[b] : mn;
    for (int b = 0; b < c->nbands; b++)
      if ((double)lin[b] > c->energy_factor * (double)mn) msk |= (1ull << b);
  }
  if (ann) { ann[0] = out[1]; ann[1] = out[2]; ann[2] = out[3]; }
  if (decision) *decision = dec;
  if (mask) *mask = msk;
}

struct sense_job {
  const crn_config *c;
  const float *iq;
  const float *win;
  int64_t g0, g1;
  float *feat;
  double *ann;
  int32_t *decision;
  uint64_t *mask;
};

static void *sense_worker(void *arg) {
  struct sense_job *j = (struct sense_job *)arg;
  const crn_config *c = j->c;
  const int N = c->nfft;
  const int stride = c->frame_stride > 0 ? c->frame_stride : c->frame_len;
  const size_t group_floats = (c->iq_format == CRN_IQ_SC16 ? 1 : 2) * (size_t)stride * c->navg; /* 4-byte units */
  cf32 *buf = (cf32 *)calloc(N, sizeof(cf32));
  cf32 *spec = (cf32 *)calloc(N, sizeof(cf32));
  float *avg = (float *)calloc(N, sizeof(float));
  fftplan plan = fft_create_plan(N, buf, spec, +1, 0); /* .cpp:42-45 */
  for (int64_t g = j->g0; g < j->g1; g++)
    sense_group(c, j->iq + g * group_floats, plan, buf, spec, avg, j->win, j->feat + g * c->nbands,
                j->ann ? j->ann + 3 * g : NULL, j->decision ? j->decision + g : NULL,
                j->mask ? j->mask + g : NULL);
  fft_destroy_plan(plan);
  free(buf);
  free(spec);
  free(avg);
  return NULL;
}

/* Returns 0, or -1 on bad arguments.  iq: ngroups*K frames, frame f at iq + 2*f*stride floats.
   Groups are independent (fft_avg is zeroed per decision, .cpp:287), so nthreads > 1 splits the group
   range into contiguous blocks, one engine state (plan + buffers) per thread. */
int crn_oracle_sense(const crn_config *c, const float *iq, int64_t ngroups, float *feat, double *ann,
                     int32_t *decision, uint64_t *mask, int nthreads) {
  if (!c || !iq || !feat || c->nfft < 2 || c->frame_len > c->nfft || c->nbands < 1) return -1;
  const int N = c->nfft;
  float *win = NULL;
  if (c->window == CRN_WINDOW_HANN) {
    win = (float *)malloc(sizeof(float) * N);
    for (int n = 0; n < N; n++) win[n] = crn_oracle_hann(n, N);
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > ngroups) nthreads = ngroups > 0 ? (int)ngroups : 1;
  struct sense_job jobs[256];
  pthread_t tid[256];
  for (int t = 0; t < nthreads; t++) {
    struct sense_job jb = {c, iq, win, ngroups * t / nthreads, ngroups * (t + 1) / nthreads,
                           feat, ann, decision, mask};
    jobs[t] = jb;
  }
  if (nthreads == 1) {
    sense_worker(&jobs[0]);
  } else {
    for (int t = 0; t < nthreads; t++) pthread_create(&tid[t], NULL, sense_worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(tid[t], NULL);
  }
  free(win);
  return 0;
}

/* Same, timed: seconds around the group loop only. */
double crn_oracle_time(const crn_config *c, const float *iq, int64_t ngroups, int nthreads) {
  float *feat = (float *)malloc(sizeof(float) * (size_t)ngroups * c->nbands);
  double *ann = (double *)malloc(sizeof(double) * 3 * (size_t)ngroups);
  int32_t *dec = (int32_t *)malloc(sizeof(int32_t) * (size_t)ngroups);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  crn_oracle_sense(c, iq, ngroups, feat, ann, dec, NULL, nthreads);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(feat);
  free(ann);
  free(dec);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int crn_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic primary-user capture (test input).  CPU statement of the generator that
 * libcrnsense's crn_synth_generate_device runs on the GPU; both follow the definition in DESIGN.md
 * ("Synthetic IQ").  It stands in for what the reference produces over the air:
 *   - PU waveform: ofdmflexframegen with the CRTS defaults - 64 subcarriers, cyclic prefix 16,
 *     taper 4 (src/crts.cpp:501-514), liquid's default subcarrier allocation (guard M/10, pilots
 *     every 8, DC null; recalled), unit power, soft gain -12 dB
 *     (src/extensible_cognitive_radio.cpp:59,892), generated at 1.4 MS/s
 *     (scenarios/predictive_model.cfg:39) and observed at 13 MS/s (:76) by evaluating the OFDM symbol
 *     as a continuous-time sum of subcarriers (an ideal resampler);
 *   - hopping: CE_PU_MARKOV_Chain_Tx.cpp:88-128 (rand()%10 outcome -> next channel) or
 *     CE_Random_Behaviour_PU.cpp:47-49 (rand()%3), one draw per dwell;
 *   - complex AWGN at the stated in-band SNR.
 * Bit-exact equality with liquid's framegen is neither possible (absent) nor needed (test input).
 * ------------------------------------------------------------------------------------------------ */
#define SY_M 64
#define SY_CP 16
#define SY_TAPER 4
#define SY_SYM (SY_M + SY_CP)
#define SY_HALF 25 /* used subcarriers k = -25..-1, 1..25 (M/2 - guard(6) - 1) */
#define SY_RATE_NUM 7
#define SY_RATE_DEN 65

static uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
uint64_t crn_oracle_mix64(uint64_t x) { return mix64(x); }

uint64_t crn_oracle_stream_seed(uint64_t seed, int64_t stream) {
  return mix64(seed ^ mix64((uint64_t)stream * 0xD1B54A32D192ED03ull + 1));
}

/* next PU channel (0..2) given the current one and an outcome r in 0..9 */
int crn_oracle_pu_next(int mode, int cur, int r) {
  if (mode == 2) return r % 3; /* CE_Random_Behaviour_PU.cpp:47-49 (r is then a raw draw) */
  if (mode == 1) return r == 0 ? 0 : 1; /* as coded: the `>=1 || <n` tests are always true (.cpp:104,114,123) */
  /* as documented (README.md:70-74; rows = current state) */
  if (r == 0) return 0;
  if (cur == 0) return r < 4 ? 1 : 2;
  if (cur == 1) return r < 6 ? 1 : 2;
  return r < 3 ? 1 : 2;
}

/* PU channel for dwells 0..ndwell-1 of one stream; dwell 0 starts on CH1 (tx_freq = 833e6,
   scenarios/predictive_model.cfg:37). */
void crn_oracle_pu_states(uint64_t stream_seed, int mode, int64_t ndwell, int8_t *out) {
  int cur = 0;
  for (int64_t d = 0; d < ndwell; d++) {
    if (d > 0) {
      const uint64_t h = mix64(stream_seed ^ (0xA5A5A5A5ull + (uint64_t)d * 0x2545F4914F6CDD1Dull));
      const int r = (mode == 2) ? (int)((h >> 33) % 3) : (int)((h >> 33) % 10);
      cur = crn_oracle_pu_next(mode, cur, r);
    }
    out[d] = (int8_t)cur;
  }
}

static void subcarrier(uint64_t sseed, int64_t m, int k, float *re, float *im) {
  /* k in -25..25, k != 0 */
  if (m < 0) { *re = *im = 0.0f; return; }
  const uint64_t h = mix64(sseed ^ mix64((uint64_t)m * 128u + (uint64_t)(k + 64)));
  const float g = 0.14142135623730950f; /* 1/sqrt(50 used subcarriers) */
  const int ak = k < 0 ? -k : k;
  if (((ak + 4) % 8) == 0) { /* pilot: BPSK */
    *re = (h & 1) ? g : -g;
    *im = 0.0f;
  } else { /* data: QPSK */
    const float a = g * 0.70710678118654752f;
    *re = (h & 1) ? a : -a;
    *im = (h & 2) ? a : -a;
  }
}

/* ---- framed waveforms (include/crnsense.h: pu_framed, CRN_INTF_GMSK / RRC / OFDM) ------------------------------- */
#define PU_FRAME_SYMS 32 /* S0 S0 S1, 7 header, 22 payload symbols (ecr.cpp:883-949 through ofdmflexframegen) */
#define OF_FRAME_SYMS 22 /* S0 S0 S1, 7 header, 12 payload symbols (interferer.cpp:255-288) */
#define OF_TAPER 6       /* interferer.hpp:24 */
#define RRC_HLEN 129     /* 2 * RRC_SAMPS_PER_SYM * RRC_FILTER_SEMILENGTH + 1 (interferer.cpp:61) */
#define RRC_FRAME 200    /* RRC_SYMS_PER_FRAME * RRC_SAMPS_PER_SYM (interferer.hpp:18-19) */
#define GM_SYMS 1024
#define GM_SPS 4         /* k = 2 (liquid gmskframegen) x the half-band interpolator (interferer.cpp:60,201-204) */
#define GM_FRAME (GM_SYMS * GM_SPS + 12) /* + zero padding (interferer.cpp:212-219) */

/* subcarrier k of the symbol at position fm of a flex frame; uid keys the random header/payload symbols */
static void frame_cell(uint64_t sseed, int64_t uid, int fm, int k, float *re, float *im) {
  *re = *im = 0.0f;
  if (uid < 0) return;
  if (fm <= 1) { /* S0: fixed sequence on the even subcarriers */
    if (k % 2 == 0) *re = (mix64(0x5330ull * 0x9E3779B97F4A7C15ull + (uint64_t)(k + 64)) & 1) ? 0.20412414523193151f : -0.20412414523193151f;
  } else if (fm == 2) { /* S1: fixed sequence on every used subcarrier */
    *re = (mix64(0x5331ull * 0x9E3779B97F4A7C15ull + (uint64_t)(k + 64)) & 1) ? 0.14142135623730950f : -0.14142135623730950f;
  } else if (fm < 10) { /* header: BPSK */
    *re = (mix64(sseed ^ mix64((uint64_t)uid * 128u + (uint64_t)(k + 64))) & 1) ? 0.14142135623730950f : -0.14142135623730950f;
  } else {
    subcarrier(sseed, uid, k, re, im);
  }
}

/* sum over the used subcarriers at phase z (z^k by recurrence), current symbol and - inside the taper - the previous one */
static void symbol_sum(uint64_t sseed, int64_t m, int frame_syms, float zr, float zi, int in_taper, float ramp,
                       float *vr, float *vi) {
  float ar = 0.0f, ai = 0.0f, br = 0.0f, bi = 0.0f, pr = zr, pi = zi;
  const int fm = frame_syms ? (int)(m % frame_syms) : 0, fp = frame_syms ? (int)((m - 1 + frame_syms) % frame_syms) : 0;
  for (int k = 1; k <= SY_HALF; k++) {
    float xr, xi, yr, yi;
    if (frame_syms) { frame_cell(sseed, m, fm, k, &xr, &xi); frame_cell(sseed, m, fm, -k, &yr, &yi); }
    else { subcarrier(sseed, m, k, &xr, &xi); subcarrier(sseed, m, -k, &yr, &yi); }
    ar += xr * pr - xi * pi + yr * pr + yi * pi;
    ai += xr * pi + xi * pr - yr * pi + yi * pr;
    if (in_taper) {
      if (frame_syms) { frame_cell(sseed, m - 1, fp, k, &xr, &xi); frame_cell(sseed, m - 1, fp, -k, &yr, &yi); }
      else { subcarrier(sseed, m - 1, k, &xr, &xi); subcarrier(sseed, m - 1, -k, &yr, &yi); }
      br += xr * pr - xi * pi + yr * pr + yi * pi;
      bi += xr * pi + xi * pr - yr * pi + yi * pr;
    }
    const float nr = pr * zr - pi * zi;
    pi = pr * zi + pi * zr;
    pr = nr;
  }
  *vr = ramp * ar + (1.0f - ramp) * br;
  *vi = ramp * ai + (1.0f - ramp) * bi;
}

/* liquid_firdes_rrcos(k = 2, m = 32, beta = 0.35, dt = 0, h) as recalled (liquid absent): tap i sits at z = i/2 - 32 symbols */
static float rrc_tap(int i) {
  const double beta = 0.35, z = (double)i / 2.0 - 32.0, g = 1.0 - 16.0 * beta * beta * z * z;
  if (fabs(z) < 1e-5) return (float)(1.0 - beta + 4.0 * beta / M_PI);
  if (fabs(g) < 1e-5)
    return (float)(beta / sqrt(2.0) * ((1.0 + 2.0 / M_PI) * sin(M_PI / (4.0 * beta)) + (1.0 - 2.0 / M_PI) * cos(M_PI / (4.0 * beta))));
  return (float)((4.0 * beta / (M_PI * g)) * (cos((1.0 + beta) * M_PI * z) + sin((1.0 - beta) * M_PI * z) / (4.0 * beta * z)));
}
/* GMSK phase pulse for BT = 0.5: a bit turns the phase by (pi/2) q(t - n - 1/2), q(x) = I(x + 1/2) - I(x - 1/2),
   I(u) = integral of the normal CDF Phi(c v) up to u = u Phi(c u) + phi(c u) / c,  c = 2 pi BT / sqrt(ln 2) */
static double gmsk_q(double x) {
  const double c = 2.0 * M_PI * 0.5 / sqrt(log(2.0));
  const double a = x + 0.5, b = x - 0.5;
  const double Ia = a * 0.5 * erfc(-c * a / sqrt(2.0)) + exp(-0.5 * c * c * a * a) / (sqrt(2.0 * M_PI) * c);
  const double Ib = b * 0.5 * erfc(-c * b / sqrt(2.0)) + exp(-0.5 * c * c * b * b) / (sqrt(2.0 * M_PI) * c);
  return Ia - Ib;
}
static uint32_t gmsk_word(uint64_t sseed, uint64_t F, int w) {
  return (uint32_t)mix64(sseed ^ mix64(0x474D534B00000000ull + F * 64ull + (uint64_t)w));
}

/* sample `im` of the interferer's own stream for the waveforms that need a modem */
static void modem_sample(int type, uint64_t sseed, uint64_t im, float *re, float *imag) {
  *re = *imag = 0.0f;
  if (type == CRN_INTF_RRC) {
    static float h[RRC_HLEN];
    static int have;
    if (!have) { for (int i = 0; i < RRC_HLEN; i++) h[i] = rrc_tap(i <= RRC_HLEN / 2 ? i : RRC_HLEN - 1 - i); have = 1; }
    const uint64_t F = im / RRC_FRAME;
    const int j = (int)(im % RRC_FRAME);
    /* firfilt reset at the frame start (:231), then y[j] = sum_i h[i] x[j-i] with a symbol on every even sample */
    for (int n = 0; 2 * n <= j; n++) {
      const int tap = j - 2 * n;
      if (tap >= RRC_HLEN) continue;
      const uint64_t hs = mix64(sseed ^ mix64(0x5252430000000000ull + F * 128ull + (uint64_t)n));
      *re = fmaf((hs & 1) ? 0.25f : -0.25f, h[tap], *re);
      *imag = fmaf((hs & 2) ? 0.25f : -0.25f, h[tap], *imag);
    }
  } else if (type == CRN_INTF_GMSK) {
    const uint64_t F = im / GM_FRAME;
    const int j = (int)(im % GM_FRAME);
    if (j >= GM_SYMS * GM_SPS) return;
    const int n0 = j / GM_SPS, sub = j % GM_SPS;
    int turns = 0; /* bits whose pulse has fully passed: +-1 quarter turn each */
    for (int n = 0; n <= n0 - 3; n++) turns += ((gmsk_word(sseed, F, n >> 5) >> (n & 31)) & 1u) ? 1 : -1;
    float frac = 0.0f;
    for (int d = -2; d <= 2; d++) {
      const int n = n0 + d;
      if (n < 0 || n >= GM_SYMS) continue;
      const float b = ((gmsk_word(sseed, F, n >> 5) >> (n & 31)) & 1u) ? 1.0f : -1.0f;
      frac = fmaf(b, (float)gmsk_q(-(double)d + (double)sub / 4.0 - 0.5), frac);
    }
    const float quarter = (float)(((turns % 4) + 4) % 4) + frac;
    *re = cosf(1.5707963267948966f * quarter);
    *imag = sinf(1.5707963267948966f * quarter);
  } else if (type == CRN_INTF_OFDM) {
    const int64_t m = (int64_t)(im / SY_SYM);
    const int tau = (int)(im % SY_SYM);
    const float th = (float)(tau - SY_CP) * (1.0f / SY_M);
    const float zr = cosf(6.283185307179586f * th), zi = sinf(6.283185307179586f * th);
    float ramp = 1.0f;
    if (tau < OF_TAPER) {
      const float sn = sinf(1.5707963267948966f * ((float)tau + 0.5f) * (1.0f / OF_TAPER));
      ramp = sn * sn;
    }
    symbol_sum(sseed ^ 0x4F46444D4F46444Dull, m, OF_FRAME_SYMS, zr, zi, tau < OF_TAPER, ramp, re, imag);
  }
}

double crn_oracle_synth_sigma2(const crn_synth_config *sc) {
  const double ps = pow(10.0, sc->pu_gain_db / 10.0);
  const double bocc = (2.0 * SY_HALF + 1.0) / SY_M * sc->pu_rate;
  return ps * sc->fs / (bocc * pow(10.0, sc->snr_db / 10.0));
}

/* One stream: samples [first, first+n) -> iq (interleaved).  states: PU channel per dwell (from
   crn_oracle_pu_states).  */
static void synth_range(const crn_synth_config *sc, uint64_t stream_seed, const int8_t *states,
                        float *iq, int64_t first, int64_t n) {
  const float gain = (float)pow(10.0, sc->pu_gain_db / 20.0);
  const float sigc = (float)sqrt(crn_oracle_synth_sigma2(sc) / 2.0);
  const int64_t dwell_samples = (int64_t)sc->dwell_groups * sc->group_samples;
  for (int64_t i = 0; i < n; i++) {
    const int64_t s = first + i;
    const int ch = states[s / dwell_samples];
    const int64_t un = s * SY_RATE_NUM;
    const int64_t m = un / ((int64_t)SY_RATE_DEN * SY_SYM);
    const float tau = (float)(un % ((int64_t)SY_RATE_DEN * SY_SYM)) / (float)SY_RATE_DEN;
    /* z = exp(j 2 pi (tau - cp)/M) */
    const float th = (tau - (float)SY_CP) * (1.0f / SY_M);
    const float zr = cosf(6.283185307179586f * th), zi = sinf(6.283185307179586f * th);
    float ramp = 1.0f;
    const int in_taper = tau < (float)SY_TAPER;
    if (in_taper) {
      const float sn = sinf(1.5707963267948966f * tau * (1.0f / SY_TAPER));
      ramp = sn * sn;
    }
    float sumr, sumi; /* current symbol, blended with the previous symbol's cyclic postfix inside the taper */
    symbol_sum(stream_seed, m, sc->pu_framed ? PU_FRAME_SYMS : 0, zr, zi, in_taper, ramp, &sumr, &sumi);
    float vr = gain * sumr;
    float vi = gain * sumi;
    /* mix to the channel offset */
    const double cyc = (double)s * (sc->offsets_hz[ch] / sc->fs);
    const float ph = (float)(cyc - floor(cyc));
    const float cr = cosf(6.283185307179586f * ph), ci = sinf(6.283185307179586f * ph);
    float outr = vr * cr - vi * ci, outi = vr * ci + vi * cr;
    /* AWGN: Box-Muller on a counter hash */
    const uint64_t h = mix64(stream_seed ^ mix64(2 * (uint64_t)s + 1));
    const float u1 = (float)((h >> 40) + 1) * (1.0f / 16777216.0f);
    const float u2 = (float)((h >> 16) & 0xFFFFFF) * (1.0f / 16777216.0f);
    const float rad = sigc * sqrtf(-2.0f * logf(u1));
    outr += rad * cosf(6.283185307179586f * u2);
    outi += rad * sinf(6.283185307179586f * u2);
    /* interferer node (src/interferer.cpp): CW :128-134, uniform noise :136-142, N(5,5) "AWGN" :24,144-154;
       its own sample rate held to ours, soft gain :32,189, mixed to its offset, duty cycle :395-409 */
    if (sc->intf_type != CRN_INTF_NONE) {
      const int64_t period = (int64_t)sc->intf_period_groups * sc->group_samples;
      const int64_t on = (int64_t)llround(sc->intf_duty * (double)period);
      if (period <= 0 || (s % period) < on) {
        const uint64_t im = (uint64_t)((double)s * (sc->intf_rate / sc->fs));
        float br2 = 0.5f, bi2 = 0.5f;
        if (sc->intf_type >= CRN_INTF_GMSK) {
          modem_sample(sc->intf_type, stream_seed, im, &br2, &bi2);
        } else if (sc->intf_type != CRN_INTF_CW) {
          const uint64_t hi = mix64(stream_seed ^ mix64(0x1F7E2A5C00000000ull + 2ull * im));
          const float v1 = (float)(hi >> 40) * (1.0f / 16777216.0f), v2 = (float)((hi >> 16) & 0xFFFFFF) * (1.0f / 16777216.0f);
          if (sc->intf_type == CRN_INTF_NOISE) {
            br2 = 0.5f * v1 - 0.25f;
            bi2 = 0.5f * v2 - 0.25f;
          } else {
            const float r5 = 5.0f * sqrtf(-2.0f * logf((float)((hi >> 40) + 1) * (1.0f / 16777216.0f)));
            br2 = 5.0f + r5 * cosf(6.283185307179586f * v2);
            bi2 = 5.0f + r5 * sinf(6.283185307179586f * v2);
          }
        }
        const double icyc = (double)s * (sc->intf_offset_hz / sc->fs);
        const float iph = (float)(icyc - floor(icyc));
        const float jr = cosf(6.283185307179586f * iph), ji = sinf(6.283185307179586f * iph);
        const float ig = (float)pow(10.0, sc->intf_gain_db / 20.0);
        outr += ig * (br2 * jr - bi2 * ji);
        outi += ig * (br2 * ji + bi2 * jr);
      }
    }
    iq[2 * i] = outr;
    iq[2 * i + 1] = outi;
  }
}

struct synth_job {
  const crn_synth_config *sc;
  uint64_t seed;
  const int8_t *states;
  float *iq;
  int64_t first, n;
};
static void *synth_worker(void *arg) {
  struct synth_job *j = (struct synth_job *)arg;
  synth_range(j->sc, j->seed, j->states, j->iq, j->first, j->n);
  return NULL;
}

/* Every sample is a pure function of its index, so the range is simply split over the host cores. */
void crn_oracle_synth(const crn_synth_config *sc, uint64_t stream_seed, const int8_t *states,
                      float *iq, int64_t first, int64_t n) {
  int nt = crn_oracle_max_threads();
  if (nt > 64) nt = 64;
  if (n < 65536 || nt < 2) {
    synth_range(sc, stream_seed, states, iq, first, n);
    return;
  }
  struct synth_job jobs[64];
  pthread_t tid[64];
  for (int t = 0; t < nt; t++) {
    const int64_t a = n * t / nt, b = n * (t + 1) / nt;
    struct synth_job jb = {sc, stream_seed, states, iq + 2 * a, first + a, b - a};
    jobs[t] = jb;
    pthread_create(&tid[t], NULL, synth_worker, &jobs[t]);
  }
  for (int t = 0; t < nt; t++) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------------------------------------------
 * The occupancy predictor on its own (checker for crn_ann_forward_device / crn_ann_train_device).
 *   forward: CE_Predictive_Node.cpp:200,214-261 on given feature rows, nothing else.
 *   train:   the reference ships only the OUTCOME of its offline training ("Error = 0.000100 after
 *            63.145737 Milion Epoch", .cpp:74; "Array of features + label", Data Generation/TODO.md:1-7) -
 *            no training code exists to restate, so this is the textbook batch back-propagation the GPU
 *            trainer documents in include/crnsense.h, written as the obvious serial loops: examples in
 *            order, E = 1/2 sum (t - Output)^2, dW = eta * (-dE/dW) / n + alpha * dW_previous.
 * ------------------------------------------------------------------------------------------------ */
static void ann_forward_one(const crn_ann_weights *w, const double *F, double *H, double *out) {
  for (int j = 1; j <= CRN_ANN_HIDDEN; j++) { /* .cpp:214-220 */
    double sum = w->wih[0][j];
    for (int i = 1; i <= CRN_ANN_INPUTS; i++) sum += F[i] * w->wih[i][j];
    H[j] = 1.0 / (1.0 + exp(-sum));
  }
  for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) { /* .cpp:229-235 */
    double sum = w->who[0][k];
    for (int j = 1; j <= CRN_ANN_HIDDEN; j++) sum += H[j] * w->who[j][k];
    out[k] = 1.0 / (1.0 + exp(-sum));
  }
}

void crn_oracle_ann_forward(const crn_ann_weights *w, double threshold, const float *feat, int64_t n,
                            int32_t stride, double *out3, int32_t *decision) {
  for (int64_t p = 0; p < n; p++) {
    const float *f = feat + p * stride;
    double F[CRN_ANN_INPUTS + 1] = {0, f[0], f[1], f[2], f[3]}; /* .cpp:200 */
    double H[CRN_ANN_HIDDEN + 1], out[CRN_ANN_OUTPUTS + 1];
    ann_forward_one(w, F, H, out);
    if (out3) { out3[3 * p] = out[1]; out3[3 * p + 1] = out[2]; out3[3 * p + 2] = out[3]; }
    if (decision) { /* .cpp:245-261 */
      int dec = CRN_ALL_BUSY;
      if (out[1] >= threshold) dec = CRN_CH1_OCCUPIED;
      else if (out[2] >= threshold) dec = CRN_CH2_OCCUPIED;
      else if (out[3] >= threshold) dec = CRN_CH3_OCCUPIED;
      decision[p] = dec;
    }
  }
}

/* E and dE/dW (as -gradient sums, i.e. the direction the update follows) at weights w for scaled inputs. */
static double ann_gradient(const crn_ann_weights *w, const double *scale, const float *feat, int32_t stride,
                           const int32_t *labels, int64_t n, crn_ann_weights *g) {
  memset(g, 0, sizeof(*g));
  double E = 0.0;
  for (int64_t p = 0; p < n; p++) {
    const float *f = feat + p * stride;
    double F[CRN_ANN_INPUTS + 1], H[CRN_ANN_HIDDEN + 1], out[CRN_ANN_OUTPUTS + 1], dO[CRN_ANN_OUTPUTS + 1];
    F[0] = 1.0;
    for (int i = 1; i <= CRN_ANN_INPUTS; i++) F[i] = (double)f[i - 1] * scale[i - 1];
    ann_forward_one(w, F, H, out);
    H[0] = 1.0;
    for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) {
      const double e = (labels[p] == k ? 1.0 : 0.0) - out[k];
      E += 0.5 * e * e;
      dO[k] = e * out[k] * (1.0 - out[k]);
    }
    for (int j = 0; j <= CRN_ANN_HIDDEN; j++)
      for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) g->who[j][k] += H[j] * dO[k];
    for (int j = 1; j <= CRN_ANN_HIDDEN; j++) {
      double sdow = 0.0;
      for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) sdow += w->who[j][k] * dO[k];
      const double dH = sdow * H[j] * (1.0 - H[j]);
      for (int i = 0; i <= CRN_ANN_INPUTS; i++) g->wih[i][j] += F[i] * dH;
    }
  }
  return E;
}

/* Error of the network at (scaled-input) weights w: used by the finite-difference gradient test. */
double crn_oracle_ann_error(const crn_ann_weights *w, const double *scale, const float *feat, int32_t stride,
                            const int32_t *labels, int64_t n, crn_ann_weights *neg_grad) {
  crn_ann_weights g;
  const double E = ann_gradient(w, scale, feat, stride, labels, n, &g);
  if (neg_grad) *neg_grad = g;
  return E;
}

int crn_oracle_ann_train(const crn_ann_train_config *tc, const float *feat, int32_t stride, const int32_t *labels,
                         int64_t n, crn_ann_weights *w, double *final_error, int32_t *epochs_run) {
  double scale[CRN_ANN_INPUTS];
  for (int i = 0; i < CRN_ANN_INPUTS; i++) scale[i] = tc->input_scale[i] != 0.0 ? tc->input_scale[i] : 1.0;
  crn_ann_weights cur, dW, g;
  memset(&cur, 0, sizeof(cur));
  memset(&dW, 0, sizeof(dW));
  if (tc->init_range > 0.0) { /* flat order: wih[i][j] -> i*5 + (j-1), then who[j][k] -> 25 + j*3 + (k-1) */
    int idx = 0;
    for (int i = 0; i <= CRN_ANN_INPUTS; i++)
      for (int j = 1; j <= CRN_ANN_HIDDEN; j++, idx++) {
        const double u = (double)(mix64(tc->seed + 0x9e3779b97f4a7c15ull * (uint64_t)(idx + 1)) >> 11) * (1.0 / 9007199254740992.0);
        cur.wih[i][j] = (2.0 * u - 1.0) * tc->init_range;
      }
    for (int j = 0; j <= CRN_ANN_HIDDEN; j++)
      for (int k = 1; k <= CRN_ANN_OUTPUTS; k++, idx++) {
        const double u = (double)(mix64(tc->seed + 0x9e3779b97f4a7c15ull * (uint64_t)(idx + 1)) >> 11) * (1.0 / 9007199254740992.0);
        cur.who[j][k] = (2.0 * u - 1.0) * tc->init_range;
      }
  } else {
    cur = *w;
    for (int i = 1; i <= CRN_ANN_INPUTS; i++)
      for (int j = 1; j <= CRN_ANN_HIDDEN; j++) cur.wih[i][j] /= scale[i - 1];
  }
  double E = 0.0;
  int epochs = 0;
  const double inv_n = 1.0 / (double)n;
  while (epochs < tc->max_epochs) {
    E = ann_gradient(&cur, scale, feat, stride, labels, n, &g);
    for (int i = 0; i <= CRN_ANN_INPUTS; i++)
      for (int j = 1; j <= CRN_ANN_HIDDEN; j++) {
        dW.wih[i][j] = tc->eta * g.wih[i][j] * inv_n + tc->alpha * dW.wih[i][j];
        cur.wih[i][j] += dW.wih[i][j];
      }
    for (int j = 0; j <= CRN_ANN_HIDDEN; j++)
      for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) {
        dW.who[j][k] = tc->eta * g.who[j][k] * inv_n + tc->alpha * dW.who[j][k];
        cur.who[j][k] += dW.who[j][k];
      }
    epochs++;
    /* the GPU trainer looks at E only every check_every epochs (and after the last one) */
    if (tc->target_error > 0.0 && E <= tc->target_error &&
        (epochs % tc->check_every == 0 || epochs == tc->max_epochs)) break;
  }
  for (int i = 1; i <= CRN_ANN_INPUTS; i++)
    for (int j = 1; j <= CRN_ANN_HIDDEN; j++) cur.wih[i][j] *= scale[i - 1];
  *w = cur;
  if (final_error) *final_error = E;
  if (epochs_run) *epochs_run = epochs;
  return 0;
}
