// Register-resident radix-R FFT codelets (R = 2..32) for sm_100a.
//
// Replaces the per-frame liquid-dsp call  fft_execute(fft)  (CE_Predictive_Node.cpp:150): forward,
// unnormalised, X[k] = sum_n x[n] exp(-j 2 pi n k / R), single precision.
//
// Design notes
//  * every index and every twiddle is a compile-time constant: after inlining the codelet is
//    straight-line FADD/FFMA code on registers, twiddles appear as FFMA immediates (the imm form
//    issues at twice the rate of the 3-register form on Blackwell, see B300_MICROARCH "Pipe rates").
//  * decimation in time on bit-reversed input, natural-order output; the bit reversal is a register
//    renaming done by the caller when it fills v[] (free).
//  * a butterfly with a non-trivial twiddle w costs 6 FFMA instead of 8 flops-as-instructions:
//        p = a + w b   (4 FFMA, the complex product folded into the accumulate)
//        q = 2a - p    (2 FFMA)
//    and 4 FADD when w is 1 or -j.
#pragma once
#include <cuda_runtime.h>

namespace crn {

// ---- compile-time trigonometry (double precision Taylor series on a reduced argument) ----------
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double cx_sin_small(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int n = 1; n < 12; n++) {
    term *= -x2 / (double)((2 * n) * (2 * n + 1));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double cx_cos_small(double x) {  // |x| <= pi/4
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int n = 1; n < 12; n++) {
    term *= -x2 / (double)((2 * n - 1) * (2 * n));
    sum += term;
  }
  return sum;
}
// cos / sin of 2 pi num / den, exact at multiples of a quarter turn.
__host__ __device__ constexpr double cx_cos_turn(int num, int den) {
  num %= den;
  if (num < 0) num += den;
  // fold to the first half turn (cos is even about 0 and about pi)
  if (2 * num > den) num = den - num;
  // now angle in [0, pi]
  if (4 * num > den) return -cx_cos_turn(den - 2 * num, 2 * den);  // cos(x) = -cos(pi - x); pi - x = 2pi (den-2num)/(2den)
  // angle in [0, pi/2]
  if (8 * num > den) {  // (pi/4, pi/2]: cos(x) = sin(pi/2 - x)
    return cx_sin_small(2.0 * kPi * (double)(den - 4 * num) / (double)(4 * den));
  }
  return cx_cos_small(2.0 * kPi * (double)num / (double)den);
}
__host__ __device__ constexpr double cx_sin_turn(int num, int den) {
  // sin(x) = cos(x - pi/2) = cos(2 pi (4 num - den) / (4 den))
  return cx_cos_turn(4 * num - den, 4 * den);
}

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
__host__ __device__ constexpr int bitrev(int v, int bits) {
  int r = 0;
  for (int b = 0; b < bits; b++) {
    r = (r << 1) | ((v >> b) & 1);
  }
  return r;
}

// ---- static loop helper ---------------------------------------------------------------------------
template <int I>
struct IC {
  static constexpr int value = I;
  constexpr operator int() const { return I; }
};
template <int B, int E, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}

// ---- butterflies ------------------------------------------------------------------------------------
// (a, b) <- (a + w b, a - w b),  w = exp(-j 2 pi J / M)
template <int M, int J>
__device__ __forceinline__ void butterfly(float2 &a, float2 &b) {
  if constexpr (J == 0) {
    const float2 p = make_float2(a.x + b.x, a.y + b.y);
    const float2 q = make_float2(a.x - b.x, a.y - b.y);
    a = p;
    b = q;
  } else if constexpr (4 * J == M) {  // w = -j : w b = (b.y, -b.x)
    const float2 p = make_float2(a.x + b.y, a.y - b.x);
    const float2 q = make_float2(a.x - b.y, a.y + b.x);
    a = p;
    b = q;
  } else {
    constexpr float c = (float)cx_cos_turn(J, M);
    constexpr float s = (float)cx_sin_turn(J, M);
    // w b = (c - j s)(bx + j by) = (c bx + s by) + j (c by - s bx)
    float pr = fmaf(c, b.x, a.x);
    pr = fmaf(s, b.y, pr);
    float pi = fmaf(c, b.y, a.y);
    pi = fmaf(-s, b.x, pi);
    b = make_float2(fmaf(2.0f, a.x, -pr), fmaf(2.0f, a.y, -pi));
    a = make_float2(pr, pi);
  }
}

// In-place radix-R FFT.  On entry v[i] = x[bitrev(i)], on exit v[k] = X[k].
// FIRST = 1 runs all log2(R) stages; FIRST = 2 skips the first (span-2, twiddle-free) stage because the
// caller has already done it, fused with a per-input weight (window or inter-pass twiddle).
template <int R, int FIRST = 1>
__device__ __forceinline__ void fft_dit(float2 (&v)[R]) {
  constexpr int LOG = ilog2(R);
  static_for<FIRST, LOG + 1>([&](auto S) {
    constexpr int m = 1 << S.value;
    constexpr int h = m >> 1;
    static_for<0, R / m>([&](auto B) {
      static_for<0, h>([&](auto J) {
        constexpr int k = B.value * m + J.value;
        butterfly<m, J.value>(v[k], v[k + h]);
      });
    });
  });
}

// First-stage butterfly fused with REAL input weights (window):  (p, q) = (wa a + wb b, wa a - wb b)
// 2 FMUL + 4 FFMA instead of 4 FMUL + 4 FADD.
__device__ __forceinline__ void butterfly_w_real(float2 a, float2 b, float wa, float wb, float2 &p, float2 &q) {
  const float ax = wa * a.x, ay = wa * a.y;
  p = make_float2(fmaf(wb, b.x, ax), fmaf(wb, b.y, ay));
  q = make_float2(fmaf(2.0f, ax, -p.x), fmaf(2.0f, ay, -p.y));
}
// ... with COMPLEX input weights (inter-pass twiddles): 2 FMUL + 8 FFMA instead of 4 FMUL + 4 FFMA + 4 FADD
// (A_IS_ONE: wa == 1, the q = 0 input of a Stockham pass: 6 FFMA).
template <bool A_IS_ONE>
__device__ __forceinline__ void butterfly_w_cplx(float2 a, float2 b, float2 wa, float2 wb, float2 &p, float2 &q) {
  float ax = a.x, ay = a.y;
  if constexpr (!A_IS_ONE) {
    ax = fmaf(a.x, wa.x, -a.y * wa.y);
    ay = fmaf(a.x, wa.y, a.y * wa.x);
  }
  float px = fmaf(b.x, wb.x, ax);
  px = fmaf(-b.y, wb.y, px);
  float py = fmaf(b.x, wb.y, ay);
  py = fmaf(b.y, wb.x, py);
  p = make_float2(px, py);
  q = make_float2(fmaf(2.0f, ax, -px), fmaf(2.0f, ay, -py));
}

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

}  // namespace crn
