// Register-resident radix-R FFT codelets (R = 2..32) for sm_100a.
//
// Replaces the per-frame liquid-dsp call  fft_execute(fft)  (CE_Predictive_Node.cpp:150): forward,
// unnormalised, X[k] = sum_n x[n] exp(-j 2 pi n k / R), single precision.
//
// Design notes
//  * every index and every twiddle is a compile-time constant: after inlining the codelet is
//    straight-line packed FADD2/FFMA2 code on 64-bit (re, im) register pairs, twiddles appear as
//    immediates.
//  * decimation in time on bit-reversed input, natural-order output; the bit reversal is a register
//    renaming done by the caller when it fills v[] (free).
//  * a butterfly with a non-trivial twiddle w costs 3 packed instructions:
//        p = a + w b   (2 FFMA2, the complex product folded into the accumulate)
//        q = 2a - p    (1 FFMA2)
//    and 2 FADD2 when w is 1 or -j; a radix-32 codelet is 194 instructions for 32 points.
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

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2) ------------------------------------------
// A complex point lives in one 64-bit register pair (re, im) - exactly what LDG.64 / LDS.64 deliver - and
// sm_100a can run an IEEE fp32 operation on both halves with ONE instruction.  The FP32 pipe still retires 32
// lane-operations per cycle per sub-partition (measured, tools/ubench_fma.cu: FFMA 0.98 and FFMA2 0.50
// warp-instructions/clk), so this does not raise the flop rate; it halves the ISSUE SLOTS the butterflies
// take, and the load/store/exchange instructions of the other warps issue in the freed slots.
// ptxas folds what the operand shapes below ask for into the instruction itself (checked in SASS):
//   make_float2(c, c) with c constant        -> 32-bit immediate, broadcast
//   make_float2(w, w) with w in a register   -> scalar-register broadcast operand  (R.F32)
//   make_float2(b.y, -b.x)                   -> half swap + sign                    (R.F32x2.LO_HI.NP)
//   make_float2(-p.x, -p.y)                  -> operand negation
// The swapped operand has to be the FIRST multiplicand and the broadcast the second, or ptxas falls back to
// FADD/MOV to build the pair.  Every half performs the same fmaf sequence as a scalar butterfly would.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float2 v) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(v.x), "f"(v.y));
  return d;
}
__device__ __forceinline__ float2 upk2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {  // a * b + c, per half
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return upk2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return upk2(d);
}
__device__ __forceinline__ float2 bc2(float w) { return make_float2(w, w); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float2 mulnegj(float2 v) { return make_float2(v.y, -v.x); }  // -j v

// ---- butterflies ------------------------------------------------------------------------------------
// (a, b) <- (a + w b, a - w b),  w = exp(-j 2 pi J / M) = c - j s:   w b = c (bx, by) + s (by, -bx)
// 3 packed instructions with a non-trivial twiddle (p = a + w b as 2 FFMA2, q = 2a - p as 1), 2 FADD2 when
// w is 1 or -j.
template <int M, int J>
__device__ __forceinline__ void butterfly(float2 &a, float2 &b) {
  if constexpr (J == 0) {
    const float2 p = add2(a, b), q = sub2(a, b);
    a = p;
    b = q;
  } else if constexpr (4 * J == M) {
    const float2 wb = mulnegj(b);
    const float2 p = add2(a, wb), q = sub2(a, wb);
    a = p;
    b = q;
  } else {
    constexpr float c = (float)cx_cos_turn(J, M);
    constexpr float s = (float)cx_sin_turn(J, M);
    float2 p = fma2(b, bc2(c), a);
    p = fma2(mulnegj(b), bc2(s), p);
    b = fma2(a, bc2(2.0f), neg2(p));
    a = p;
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

// Butterfly with a run-time twiddle tau = (c, -s) held in registers: 3 packed instructions.
__device__ __forceinline__ void butterfly_rt(float2 &a, float2 &b, float c, float s) {
  float2 p = fma2(b, bc2(c), a);
  p = fma2(mulnegj(b), bc2(s), p);
  b = fma2(a, bc2(2.0f), neg2(p));
  a = p;
}

// "Twisted" radix-R DIT codelet: the inter-pass twiddles w^q of a Stockham pass (input q of the column is
// multiplied by w^q, w a per-thread constant) are not applied to the inputs; they ride on the butterflies:
//     X[k] = sum_q x_q w^q W_R^(qk)   =>   stage m = 2^S multiplies by  tau_m[J] = w^(R/m) W_m^J  instead of W_m^J.
// Every butterfly then has a run-time twiddle (3 packed instructions, none trivial), but there is no separate
// twiddle layer: a radix-32 pass costs 240 packed instructions against 194 + 62 (+ the table reads of 31 twiddles).
// tau_m[J + m/4] = -j tau_m[J], so a stage needs max(1, m/4) table values: R/2 in all, laid out stage by stage
//     entry 0: tau_2[0];  entry m/4 + J: tau_m[J], J < m/4  (m >= 4),
// two entries per 16-byte table row:  col[(e/2) * RS] = { tau_e, tau_(e+1) }  as (re, im, re, im).
// On entry v[i] = x[bitrev(i)]; FIRST = 2 skips stage 1 (done by the caller).
template <int RS>
struct TwistedTable {
  const float4 *__restrict__ col;  // this thread's column of the table; rows are RS float4 apart
  template <int E>
  __device__ __forceinline__ float2 entry() const {
    const float4 w = col[(E / 2) * RS];
    return (E & 1) ? make_float2(w.z, w.w) : make_float2(w.x, w.y);
  }
};
template <int R, int FIRST, int RS>
__device__ __forceinline__ void fft_dit_twisted(float2 (&v)[R], const TwistedTable<RS> tw) {
  constexpr int LOG = ilog2(R);
  static_for<FIRST, LOG + 1>([&](auto S) {
    constexpr int m = 1 << S.value;
    constexpr int h = m >> 1;
    constexpr int cnt = (m >= 4) ? m / 4 : 1;
    constexpr int base = (m >= 4) ? m / 4 : 0;
    static_for<0, cnt>([&](auto J) {
      const float2 tau = tw.template entry<base + J.value>();
      static_for<0, R / m>([&](auto B) {
        constexpr int k = B.value * m + J.value;
        butterfly_rt(v[k], v[k + h], tau.x, -tau.y);
        if constexpr (m >= 4) butterfly_rt(v[k + cnt], v[k + cnt + h], tau.y, tau.x);  // -j tau
      });
    });
  });
}

// a * w (complex): 2 packed instructions
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return fma2(mulnegj(a), bc2(-w.y), mul2(a, bc2(w.x)));
}
// acc + b * w (complex): 2 packed instructions
__device__ __forceinline__ float2 cmadd(float2 b, float2 w, float2 acc) {
  return fma2(mulnegj(b), bc2(-w.y), fma2(b, bc2(w.x), acc));
}

// First-stage butterfly fused with REAL input weights (window):  (p, q) = (wa a + wb b, wa a - wb b)
// 3 packed instructions.
__device__ __forceinline__ void butterfly_w_real(float2 a, float2 b, float wa, float wb, float2 &p, float2 &q) {
  const float2 A = mul2(a, bc2(wa));
  p = fma2(b, bc2(wb), A);
  q = fma2(A, bc2(2.0f), neg2(p));
}
// ... with COMPLEX input weights (inter-pass twiddles): 5 packed instructions
// (A_IS_ONE: wa == 1, the q = 0 input of a Stockham pass: 3).
template <bool A_IS_ONE>
__device__ __forceinline__ void butterfly_w_cplx(float2 a, float2 b, float2 wa, float2 wb, float2 &p, float2 &q) {
  float2 A = a;
  if constexpr (!A_IS_ONE) A = cmul(a, wa);
  p = cmadd(b, wb, A);
  q = fma2(A, bc2(2.0f), neg2(p));
}

}  // namespace crn
