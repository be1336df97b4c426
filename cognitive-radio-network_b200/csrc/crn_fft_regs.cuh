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
  template <int ROW>
  __device__ __forceinline__ float4 row() const { return col[ROW * RS]; }
};

// ---- tensor memory (TMEM) as a per-thread table store ----------------------------------------------------------
// Blackwell's 256 KB of tensor memory per SM (512 columns x 128 lanes x 32 bit) is idle in a kernel without MMAs, and
// tcgen05.ld / tcgen05.st move it to and from registers through their own pipe, not the L1 / shared-memory data pipe
// that bounds the hybrid FFT plans.  With the .32x32b shape thread i of a warp owns lane 32 (warp % 4) + i and reads
// consecutive columns: a private, loop-invariant table of the thread - exactly what the twisted codelets' twiddle
// columns are.  A thread only ever reads what it wrote itself, so no cross-thread ordering is involved.
__device__ __forceinline__ void tmem_st4(unsigned taddr, float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
               "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
               : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A tcgen05.ld in flight: the registers it will fill.  tmem_wait hands them out - the wait instruction carries them as
// in/out operands, so nothing can consume them early, while independent arithmetic is free to move in between.
struct TmemPending4 { unsigned x, y, z, w; };
struct TmemPending8 { unsigned r[8]; };
__device__ __forceinline__ TmemPending4 tmem_issue4(unsigned taddr) {
  TmemPending4 p;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(p.x), "=r"(p.y), "=r"(p.z), "=r"(p.w) : "r"(taddr));
  return p;
}
__device__ __forceinline__ float4 tmem_wait(TmemPending4 p) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(p.x), "+r"(p.y), "+r"(p.z), "+r"(p.w));
  return make_float4(__uint_as_float(p.x), __uint_as_float(p.y), __uint_as_float(p.z), __uint_as_float(p.w));
}
__device__ __forceinline__ TmemPending8 tmem_issue8(unsigned taddr) {
  TmemPending8 p;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(p.r[0]), "=r"(p.r[1]), "=r"(p.r[2]), "=r"(p.r[3]), "=r"(p.r[4]), "=r"(p.r[5]), "=r"(p.r[6]), "=r"(p.r[7])
               : "r"(taddr));
  return p;
}
__device__ __forceinline__ void tmem_wait(TmemPending8 p, float4 &a, float4 &b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(p.r[0]), "+r"(p.r[1]), "+r"(p.r[2]), "+r"(p.r[3]), "+r"(p.r[4]), "+r"(p.r[5]), "+r"(p.r[6]), "+r"(p.r[7]));
  a = make_float4(__uint_as_float(p.r[0]), __uint_as_float(p.r[1]), __uint_as_float(p.r[2]), __uint_as_float(p.r[3]));
  b = make_float4(__uint_as_float(p.r[4]), __uint_as_float(p.r[5]), __uint_as_float(p.r[6]), __uint_as_float(p.r[7]));
}
__device__ __forceinline__ void tmem_wait(TmemPending8 p, float (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(p.r[0]), "+r"(p.r[1]), "+r"(p.r[2]), "+r"(p.r[3]), "+r"(p.r[4]), "+r"(p.r[5]), "+r"(p.r[6]), "+r"(p.r[7]));
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(p.r[i]);
}
__device__ __forceinline__ float4 tmem_ld4(unsigned taddr) { return tmem_wait(tmem_issue4(taddr)); }
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float4 &a, float4 &b) { tmem_wait(tmem_issue8(taddr), a, b); }
struct TmemTwistedTable {
  unsigned taddr;  // TMEM address (lane base of the warp | first column) of this thread's rows, 4 columns per row
  template <int ROW>
  __device__ __forceinline__ float4 row() const { return tmem_ld4(taddr + 4 * ROW); }
};

template <int R, int FIRST, class Table>
__device__ __forceinline__ void fft_dit_twisted(float2 (&v)[R], const Table tw) {
  constexpr int LOG = ilog2(R);
  static_for<FIRST, LOG + 1>([&](auto S) {
    constexpr int m = 1 << S.value;
    constexpr int h = m >> 1;
    constexpr int cnt = (m >= 4) ? m / 4 : 1;
    constexpr int base = (m >= 4) ? m / 4 : 0;
    // entries base .. base + cnt - 1, two per table row
    static_for<0, (cnt + 1) / 2>([&](auto JR) {
      const float4 rw = tw.template row<(base + 2 * JR.value) / 2>();
      static_for<0, (cnt >= 2 ? 2 : 1)>([&](auto H) {
        constexpr int e = base + 2 * JR.value + H.value;
        constexpr int J = e - base;
        const float2 tau = (e & 1) ? make_float2(rw.z, rw.w) : make_float2(rw.x, rw.y);
        static_for<0, R / m>([&](auto B) {
          constexpr int k = B.value * m + J;
          butterfly_rt(v[k], v[k + h], tau.x, -tau.y);
          if constexpr (m >= 4) butterfly_rt(v[k + cnt], v[k + cnt + h], tau.y, tau.x);  // -j tau
        });
      });
    });
  });
}

// The whole twisted radix-R codelet (stage 1 included) with the table rows fetched from tensor memory one row ahead:
// row r + 1 is in flight while the butterflies of row r run (a row feeds 4 .. R/2 butterflies, the load takes ~12
// cycles).  Entry e of the table belongs to stage m = 2 (e = 0) or m = 4 << floor(log2 e), J = e - m/4.
template <int R>
__device__ __forceinline__ void fft_dit_twisted_tmem(float2 (&v)[R], unsigned taddr) {
  TmemPending4 pend = tmem_issue4(taddr);
  static_for<0, R / 4>([&](auto ROW) {
    const float4 rw = tmem_wait(pend);
    if constexpr (ROW.value + 1 < R / 4) pend = tmem_issue4(taddr + 4 * (ROW.value + 1));
    static_for<0, 2>([&](auto H) {
      constexpr int e = 2 * ROW.value + H.value;
      const float2 tau = H.value ? make_float2(rw.z, rw.w) : make_float2(rw.x, rw.y);
      if constexpr (e == 0) {
        static_for<0, R / 2>([&](auto B) { butterfly_rt(v[2 * B.value], v[2 * B.value + 1], tau.x, -tau.y); });
      } else {
        constexpr int m = 4 << ilog2(e), h = m / 2, cnt = m / 4, J = e - cnt;
        static_for<0, R / m>([&](auto B) {
          constexpr int k = B.value * m + J;
          butterfly_rt(v[k], v[k + h], tau.x, -tau.y);
          butterfly_rt(v[k + cnt], v[k + cnt + h], tau.y, tau.x);  // -j tau
        });
      }
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
