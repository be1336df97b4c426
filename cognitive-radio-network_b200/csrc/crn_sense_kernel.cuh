// Fused spectrum-sensing kernel for sm_100a.
//
// One launch does, for every decision group (K frames of L <= N samples), what the reference engine
// does across K calls of CE_Predictive_Node::execute():
//     .cpp:149      stage L samples, zero tail                  -> predicated 8-byte global loads
//     (extension)   window multiply                             -> smem table, fused into the load
//     .cpp:150      fft_execute (N-point forward FFT)           -> register radix-16/32 passes with at
//                                                                  most two shared-memory exchanges
//     .cpp:152-154  fft_avg[i] += |X[i]| / K   (or |X[i]|^2)    -> per-thread register accumulators
//     .cpp:173-191  band sums                                    -> warp-shuffle segmented reduction
//     .cpp:194-200  CHx = Mx*Mx, feature vector                  -> fused
//     .cpp:214-235  4-5-3 logistic MLP in double                 -> one lane, FMA + exp in fp64
//     .cpp:245-261  first-match threshold chain                  -> decision code
// IQ is read from HBM exactly once (8 B/sample); only nbands floats + 3 doubles + 12 bytes per group
// are written back.
//
// FFT decomposition (Stockham autosort, decimation in time): N = R0*R1(*R2).  A frame is owned by a
// "team" of T = N/E threads, each holding E complex points in registers.  In every pass thread t holds
// the points {t + T*m, m = 0..E-1}: the first pass therefore reads global memory fully coalesced
// (consecutive lanes -> consecutive samples), every exchange reads shared memory at unit stride, and
// the last pass leaves bin (t + T*m) in register m, so the K-frame accumulators never leave registers.
#pragma once
#include <cstdint>

#include "crn_fft_regs.cuh"
#include "crnsense.h"

namespace crn {

enum { DET_MAG = 0, DET_MAGSQ = 1 };

struct SenseParams {
  const float2 *iq;      // [ngroups][K][stride] complex-float
  const float2 *tw;      // inter-pass twiddles: pass-1 table then pass-2 table
  const float *win;      // [N] window (nullptr for rectangular)
  float *feat;           // [ngroups][nbands]
  double *ann;           // [ngroups][3] or nullptr
  int32_t *decision;     // [ngroups] or nullptr
  unsigned long long *mask;  // [ngroups] or nullptr
  long long ngroups;
  int L, stride, K;
  float invK;
  int nbands, nsegs, postop, decide;
  double threshold, energy_factor;
  double wih[CRN_ANN_INPUTS + 1][CRN_ANN_HIDDEN + 1];
  double who[CRN_ANN_HIDDEN + 1][CRN_ANN_OUTPUTS + 1];
  short seg_band[CRN_MAX_SEGS], seg_lo[CRN_MAX_SEGS], seg_hi[CRN_MAX_SEGS];
};

// Compile-time plan for one FFT size.
template <int N_, int E_, int R0_, int R1_, int R2_, int TEAMS_, int MINB_>
struct Plan {
  static constexpr int N = N_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_, TEAMS = TEAMS_, MINB = MINB_;
  static constexpr int T = N / E;                  // threads per frame
  static constexpr int NT = T * TEAMS;             // threads per CTA
  static constexpr int PASSES = (R2 > 1) ? 3 : 2;
  static constexpr int PADSHIFT = ilog2(R0);       // one pad slot per R0 points: conflict-free exchange
  static constexpr int XSZ = N + (N >> PADSHIFT);  // float2 slots per team exchange buffer
  static constexpr int TW1 = R0 * R1;              // pass-1 twiddle table entries
  static constexpr int TW2 = (R2 > 1) ? N : 0;     // pass-2 twiddle table entries
  static_assert(R0 * R1 * R2 == N, "radices must multiply to N");
  static_assert(E % R0 == 0 && E % R1 == 0 && E % R2 == 0, "E must be a multiple of every radix");
  static_assert(R0 >= 16, "first radix < 16 would bank-conflict the exchange");
  static_assert(T >= 16 && (T <= 32 ? 32 % T == 0 : T % 32 == 0), "team must tile a warp");
  static constexpr size_t smem_bytes(bool win) {
    return sizeof(float2) * ((size_t)TEAMS * XSZ + TW1 + TW2) + (win ? sizeof(float) * N : 0) +
           sizeof(float) * (CRN_MAX_SEGS + CRN_MAX_BANDS);
  }
};

__device__ __forceinline__ float2 ld_stream(const float2 *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

template <int T>
__device__ __forceinline__ void team_sync(int team) {
  if constexpr (T == 32) {
    __syncwarp();
  } else if constexpr (T < 32) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned base = lane & ~(unsigned)(T - 1);
    __syncwarp((T == 16 ? 0xFFFFu : 0xFFu) << base);
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(T) : "memory");
  }
}

// One radix-R pass over the E register-resident points: E/R independent FFTs, FFT #i on registers
// {i + r*(E/R)}.
template <int E, int R>
__device__ __forceinline__ void reg_pass(float2 (&a)[E]) {
  constexpr int G = E / R;
  constexpr int LOG = ilog2(R);
  static_for<0, G>([&](auto I) {
    float2 v[R];
    static_for<0, R>([&](auto Q) {
      constexpr int br = bitrev(Q.value, LOG);
      v[br] = a[I.value + Q.value * G];
    });
    fft_dit<R>(v);
    static_for<0, R>([&](auto Q) { a[I.value + Q.value * G] = v[Q.value]; });
  });
}

// Multiply by the inter-pass twiddles W_{Ns*R}^{r*(j mod Ns)}, j = t + T*i.
template <int E, int R, int T, int NS>
__device__ __forceinline__ void apply_twiddles(float2 (&a)[E], const float2 *__restrict__ tw, int t) {
  constexpr int G = E / R;
  static_for<0, G>([&](auto I) {
    const int q = (t + T * I.value) & (NS - 1);
    static_for<1, R>([&](auto Q) {
      const float2 w = tw[Q.value * NS + q];
      a[I.value + Q.value * G] = cmul(a[I.value + Q.value * G], w);
    });
  });
}

// Scatter the outputs of a radix-R pass (Ns = product of earlier radices) into the exchange buffer,
// then gather this thread's points for the next pass.
template <int E, int R, int T, int NS, int PADSHIFT>
__device__ __forceinline__ void exchange(float2 (&a)[E], float2 *__restrict__ xb, int t, int team) {
  constexpr int G = E / R;
  team_sync<T>(team);  // previous readers of xb are done
  static_for<0, G>([&](auto I) {
    const int j = t + T * I.value;
    const int base = (j / NS) * (NS * R) + (j & (NS - 1));
    static_for<0, R>([&](auto Q) {
      const int idx = base + Q.value * NS;
      xb[idx + (idx >> PADSHIFT)] = a[I.value + Q.value * G];
    });
  });
  team_sync<T>(team);
  static_for<0, E>([&](auto M) {
    const int idx = t + T * M.value;
    a[M.value] = xb[idx + (idx >> PADSHIFT)];
  });
}

template <class P, bool WIN, int DET>
__global__ void __launch_bounds__(P::NT, P::MINB) sense_kernel(const SenseParams prm) {
  constexpr int N = P::N, E = P::E, T = P::T, TEAMS = P::TEAMS, NT = P::NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *xbuf = reinterpret_cast<float2 *>(smem_raw);
  float2 *tw1 = xbuf + (size_t)TEAMS * P::XSZ;
  float2 *tw2 = tw1 + P::TW1;
  float *fl = reinterpret_cast<float *>(tw2 + P::TW2);
  float *win = fl;
  float *segsum = fl + (WIN ? N : 0);
  float *featbuf = segsum + CRN_MAX_SEGS;

  const int tid = threadIdx.x;
  const int team = tid / T;
  const int t = tid % T;
  float2 *xb = xbuf + (size_t)team * P::XSZ;

  // one-time table staging (persistent CTA: amortised over all its groups)
  for (int i = tid; i < P::TW1 + P::TW2; i += NT) tw1[i] = prm.tw[i];
  if constexpr (WIN)
    for (int i = tid; i < N; i += NT) win[i] = prm.win[i];
  __syncthreads();

  const int L = prm.L, K = prm.K;
  const bool full = (L == N);

  for (long long g = blockIdx.x; g < prm.ngroups; g += gridDim.x) {
    float acc[E];
#pragma unroll
    for (int m = 0; m < E; m++) acc[m] = 0.0f;

    const float2 *gbase = prm.iq + (size_t)g * (size_t)K * (size_t)prm.stride;
    for (int k = team; k < K; k += TEAMS) {
      const float2 *x = gbase + (size_t)k * (size_t)prm.stride + t;
      float2 a[E];
      if (full) {
#pragma unroll
        for (int m = 0; m < E; m++) a[m] = ld_stream(x + T * m);
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) a[m] = (t + T * m < L) ? ld_stream(x + T * m) : make_float2(0.f, 0.f);
      }
      if constexpr (WIN) {
#pragma unroll
        for (int m = 0; m < E; m++) {
          const float w = win[t + T * m];
          a[m].x *= w;
          a[m].y *= w;
        }
      }
      // pass 0 (no twiddles: Ns = 1)
      reg_pass<E, P::R0>(a);
      exchange<E, P::R0, T, 1, P::PADSHIFT>(a, xb, t, team);
      // pass 1
      apply_twiddles<E, P::R1, T, P::R0>(a, tw1, t);
      reg_pass<E, P::R1>(a);
      if constexpr (P::PASSES == 3) {
        exchange<E, P::R1, T, P::R0, P::PADSHIFT>(a, xb, t, team);
        apply_twiddles<E, P::R2, T, P::R0 * P::R1>(a, tw2, t);
        reg_pass<E, P::R2>(a);
      }
      // register m now holds bin t + T*m  (.cpp:152-154)
#pragma unroll
      for (int m = 0; m < E; m++) {
        if constexpr (DET == DET_MAGSQ) {
          acc[m] = fmaf(a[m].x, a[m].x, acc[m]);
          acc[m] = fmaf(a[m].y, a[m].y, acc[m]);
        } else {
          float s;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(a[m].x, a[m].x, a[m].y * a[m].y)));
          acc[m] += s;
        }
      }
    }

    // ---- per-group epilogue: band sums, features, MLP, decision ------------------------------------
    __syncthreads();  // every team finished reading its exchange buffer
    {
      float *part = reinterpret_cast<float *>(xb);  // N floats per team, aliasing the exchange buffer
#pragma unroll
      for (int m = 0; m < E; m++) part[t + T * m] = acc[m];
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      constexpr int NW = NT / 32;
      for (int s = warp; s < prm.nsegs; s += NW) {
        float sum = 0.0f;
        for (int i = prm.seg_lo[s] + lane; i < prm.seg_hi[s]; i += 32) {
#pragma unroll
          for (int q = 0; q < TEAMS; q++) sum += reinterpret_cast<const float *>(xbuf + (size_t)q * P::XSZ)[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) segsum[s] = sum;
      }
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int b = lane; b < prm.nbands; b += 32) {
        float m = 0.0f;
        for (int s = 0; s < prm.nsegs; s++)
          if (prm.seg_band[s] == b) m += segsum[s];
        m *= prm.invK;
        const float f = (prm.postop == CRN_POST_SQUARE_OF_SUM) ? m * m : m;  // .cpp:194-197
        featbuf[b] = f;
        prm.feat[(size_t)g * prm.nbands + b] = f;
      }
      __syncwarp();
      if (prm.decide == CRN_DECIDE_ANN) {
        if (lane == 0) {
          // .cpp:200,214-235: double-precision 4-5-3 logistic MLP, bias at index 0
          double H[CRN_ANN_HIDDEN + 1];
#pragma unroll
          for (int j = 1; j <= CRN_ANN_HIDDEN; j++) {
            double sum = prm.wih[0][j];
#pragma unroll
            for (int i = 1; i <= CRN_ANN_INPUTS; i++) sum += (double)featbuf[i - 1] * prm.wih[i][j];
            H[j] = 1.0 / (1.0 + exp(-sum));
          }
          double out[CRN_ANN_OUTPUTS + 1];
#pragma unroll
          for (int k = 1; k <= CRN_ANN_OUTPUTS; k++) {
            double sum = prm.who[0][k];
#pragma unroll
            for (int j = 1; j <= CRN_ANN_HIDDEN; j++) sum += H[j] * prm.who[j][k];
            out[k] = 1.0 / (1.0 + exp(-sum));
          }
          int dec = CRN_ALL_BUSY;  // .cpp:245-261
          if (out[1] >= prm.threshold) dec = CRN_CH1_OCCUPIED;
          else if (out[2] >= prm.threshold) dec = CRN_CH2_OCCUPIED;
          else if (out[3] >= prm.threshold) dec = CRN_CH3_OCCUPIED;
          if (prm.ann) {
            prm.ann[3 * g + 0] = out[1];
            prm.ann[3 * g + 1] = out[2];
            prm.ann[3 * g + 2] = out[3];
          }
          if (prm.decision) prm.decision[g] = dec;
          if (prm.mask) prm.mask[g] = dec ? (1ull << (dec - 1)) : 0ull;
        }
      } else {
        unsigned long long msk = 0ull;
        if (prm.decide == CRN_DECIDE_ENERGY) {
          float mn = 3.4e38f;
          for (int b = lane; b < prm.nbands; b += 32) mn = fminf(mn, featbuf[b]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          for (int b = lane; b < prm.nbands; b += 32)
            if ((double)featbuf[b] > prm.energy_factor * (double)mn) msk |= (1ull << b);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) msk |= __shfl_xor_sync(0xffffffffu, msk, o);
        }
        if (lane == 0) {
          if (prm.ann) {
            prm.ann[3 * g + 0] = 0.0;
            prm.ann[3 * g + 1] = 0.0;
            prm.ann[3 * g + 2] = 0.0;
          }
          if (prm.decision) prm.decision[g] = 0;
          if (prm.mask) prm.mask[g] = msk;
        }
      }
    }
    // The next group's first exchange starts with a team_sync, but other warps may still be reading
    // `part` in the band reduction of THIS group only before the barrier above - nothing after it
    // reads the exchange buffers, so no further barrier is needed here.
  }
}

}  // namespace crn
