/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into libcrnsense or any product path.
 *
 * Restatement of the two liquid-dsp entry points the reference engine calls:
 *     fft_create_plan()  CE_Predictive_Node.cpp:42-45
 *     fft_execute()      CE_Predictive_Node.cpp:150
 * liquid-dsp is a third-party dependency that is ABSENT from /root/reference (pinned at git a4d7c80d3
 * by HardwareSetup/Install_liquid-dsp.sh:36; its only local patch touches poly.findroots.c).  Its
 * published algorithm for power-of-two sizes (src/fft/src/fft_radix2.c, recalled, source not
 * available here) is: single-precision, out of place, unnormalised forward transform
 * X[k] = sum_n x[n] exp(-j 2 pi n k / N); the input is copied to the output in bit-reversed order and
 * log2(N) in-place decimation-in-time radix-2 stages follow, the stage twiddle taken from an N-entry
 * table exp(-j 2 pi i / N) (built with cexpf) that is walked with stride N/(2*half) for butterfly
 * column j.  A liquid built on FFTW3 would dispatch to fftwf instead (the ECR carries a CE_fftw_mutex,
 * include/extensible_cognitive_radio.hpp:880-884); both agree to single-precision rounding.
 * => PARITY UNPINNED by the reference (it ships no test or golden vector for this boundary).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>

typedef float _Complex cf32;

struct fftplan_s {
  unsigned int n, log2n;
  int dir;
  cf32 *x, *y;
  cf32 *tw;           /* tw[i] = exp(dir * -j 2 pi i / n) */
  unsigned int *rev;  /* bit-reversed index */
};
typedef struct fftplan_s *fftplan;

static unsigned int bitrev(unsigned int v, unsigned int bits) {
  unsigned int r = 0;
  for (unsigned int b = 0; b < bits; b++) {
    r = (r << 1) | (v & 1u);
    v >>= 1;
  }
  return r;
}

fftplan fft_create_plan(unsigned int n, cf32 *x, cf32 *y, int dir, int flags) {
  (void)flags;
  fftplan p = (fftplan)calloc(1, sizeof(*p));
  if (!p) return NULL;
  p->n = n;
  p->x = x;
  p->y = y;
  p->dir = dir;
  unsigned int lg = 0;
  while ((1u << lg) < n) lg++;
  p->log2n = lg;
  p->tw = (cf32 *)malloc(sizeof(cf32) * n);
  p->rev = (unsigned int *)malloc(sizeof(unsigned int) * n);
  const double sgn = (dir > 0) ? -1.0 : 1.0; /* LIQUID_FFT_FORWARD == +1 */
  for (unsigned int i = 0; i < n; i++) {
    /* the angle is formed in double (M_PI is a double) and handed to the float exponential */
    float _Complex arg = (float _Complex)(_Complex_I * (sgn * 2.0 * M_PI * (double)i / (double)n));
    p->tw[i] = cexpf(arg);
    p->rev[i] = ((1u << lg) == n) ? bitrev(i, lg) : i;
  }
  return p;
}

void fft_destroy_plan(fftplan p) {
  if (!p) return;
  free(p->tw);
  free(p->rev);
  free(p);
}

static void dft_direct(fftplan p) {
  /* not a power of two: plain O(n^2) evaluation (never hit by the reference's n = 512) */
  const double sgn = (p->dir > 0) ? -1.0 : 1.0;
  for (unsigned int k = 0; k < p->n; k++) {
    double _Complex acc = 0;
    for (unsigned int i = 0; i < p->n; i++)
      acc += (double _Complex)p->x[i] *
             cexp(_Complex_I * (sgn * 2.0 * M_PI * (double)((unsigned long long)i * k % p->n) / p->n));
    p->y[k] = (cf32)acc;
  }
}

void fft_execute(fftplan p) {
  const unsigned int n = p->n;
  if ((1u << p->log2n) != n) {
    dft_direct(p);
    return;
  }
  cf32 *y = p->y;
  for (unsigned int i = 0; i < n; i++) y[i] = p->x[p->rev[i]];
  unsigned int half = 1, stride = n;
  for (unsigned int s = 0; s < p->log2n; s++) {
    const unsigned int span = half * 2;
    stride >>= 1;
    unsigned int ti = 0;
    for (unsigned int j = 0; j < half; j++) {
      const cf32 w = p->tw[ti];
      ti = (ti + stride) % n;
      for (unsigned int k = j; k < n; k += span) {
        const cf32 t = y[k + half] * w;
        y[k + half] = y[k] - t;
        y[k] += t;
      }
    }
    half = span;
  }
}
