// Fused spectrum-sensing kernel for sm_100a.
//
// One launch does, for every decision group (K frames of L <= N samples), what the reference engine
// does across K calls of CE_Predictive_Node::execute():
//     .cpp:149      stage L samples, zero tail                  -> predicated 8-byte global loads
//     (extension)   window multiply                             -> folded into the first butterfly stage
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
// The two inputs of every first-stage butterfly are registers (m, m + E/2), whatever the radix, so the
// per-input weights (window in pass 0, inter-pass twiddles later) are stored as pairs and folded into
// that stage.
//
// Work distribution.  A CTA is NT threads = UNITS reduction units (a unit is a warp, or a whole team
// when a team spans several warps).  `upg` units share one decision group (its K frames are dealt
// round-robin to their teams), so a CTA works on UNITS/upg groups at a time.  There is no CTA-wide
// barrier in the steady state: each unit reduces its own accumulators to per-segment partial sums, and
// the last unit of a group to arrive (shared-memory counter) combines them, runs the MLP and writes the
// decision while the others already work on their next group.
//
// Group splitting (CTA epilogue only).  With few groups per launch - one decision on the streaming path, a tail
// round that fills half the GPU - a whole group per CTA leaves SMs idle.  The host may therefore deal a group's
// K frames to `split` work items of K/split consecutive frames; each item leaves its per-segment sums in a
// scratch row, and the item that arrives last (one atomic per item, the classic threadfence reduction) adds the
// rows in part order and takes the decision.  The order of additions is fixed by (K, split), never by timing.
#pragma once
#include <cstdint>
#include <type_traits>

#include "crn_fft_regs.cuh"
#include "crnsense.h"

namespace crn {

enum { DET_MAG = 0, DET_MAGSQ = 1 };

struct SenseParams {
  const float2 *iq;      // [ngroups][K][stride] complex-float
  const float4 *tw;      // paired inter-pass twiddles: pass-1 table then pass-2 table
  const float2 *winp;    // [N/2] paired window (nullptr for rectangular)
  float *feat;           // [ngroups][nbands]
  double *ann;           // [ngroups][3] or nullptr
  int32_t *decision;     // [ngroups] or nullptr
  unsigned long long *mask;  // [ngroups] or nullptr
  long long ngroups;
  int L, stride, K;
  float invK;
  int nbands, nsegs, postop, decide;
  int upg;               // reduction units per decision group (divides UNITS); 0 = CTA-wide epilogue
  int use_tma;           // 1: stage frames with cp.async.bulk (needs 16-byte aligned frames, CTA epilogue)
  int split;             // CTA epilogue: a group's K frames are dealt to `split` work items (K % (split*TEAMS) == 0)
  int kp;                // K / split: frames per work item
  long long nwork;       // ngroups * split
  float *scratch;        // [ngroups][split][nsegs] per-part segment sums (split > 1)
  int *gcount;           // [ngroups] parts arrived, zero between launches (split > 1)
  int sc16;              // IQ buffers hold int16 pairs (4 B/sample) instead of float pairs
  int seg_stride, band_stride;  // row lengths of the epilogue's shared-memory scratch: nsegs / nbands rounded up to 8
  unsigned acc_mask;     // bit m set: some band segment reads a bin held in accumulator register m
  double threshold, energy_factor;
  double wih[CRN_ANN_INPUTS + 1][CRN_ANN_HIDDEN + 1];
  double who[CRN_ANN_HIDDEN + 1][CRN_ANN_OUTPUTS + 1];
  short seg_band[CRN_MAX_SEGS], seg_lo[CRN_MAX_SEGS], seg_hi[CRN_MAX_SEGS];
  // Band b's segments are entries [band_first[b], band_first[b + 1]) when the table lists them band by band
  // (bands_contig; every built-in plan does): the combine then walks its own segments instead of scanning all nsegs
  // for every band (64 x 64 tests on one warp per decision in the 64-sub-channel plans).  Same additions, same order.
  short band_first[CRN_MAX_BANDS + 1];
  int bands_contig;
};

// Compile-time plan for one FFT size.
template <int N_, int E_, int R0_, int R1_, int R2_, int TEAMS_, int MINB_>
struct Plan {
  static constexpr int N = N_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_, TEAMS = TEAMS_, MINB = MINB_;
  static constexpr int T = N / E;                  // threads per frame
  static constexpr int NT = T * TEAMS;             // threads per CTA
  static constexpr int PASSES = (R2 > 1) ? 3 : 2;
  static constexpr int PADSHIFT = ilog2(R0);       // two pad slots per R0 points: conflict-free, 16 B rows
  static constexpr int XSZ = N + 2 * (N >> PADSHIFT);  // float2 slots per team exchange buffer
  static constexpr int TW1 = R0 * R1 / 4;          // pass-1 twisted-codelet table: R1/4 rows of R0 columns (float4)
  static constexpr int TW2 = (R2 > 1) ? N / 4 : 0; // pass-2 table: R2/4 rows of R0*R1 columns
  static constexpr int UNIT_THREADS = T > 32 ? T : 32;
  static constexpr int UNITS = NT / UNIT_THREADS;  // reduction units per CTA
  static constexpr int TEAMS_PER_UNIT = UNIT_THREADS / T;
  static constexpr bool HYBRID = false;
  static constexpr bool EARLY_FREE = false;
  static constexpr int C = 1;
  // Pass-1 twiddle columns, window pairs and the all-bins accumulators in tensor memory, as in the hybrid plans (see
  // HybridPlan::TMEM_TW): ncu puts the L1 / shared-memory data pipe of the N = 1024 kernel at 82 % busy with HBM at
  // the copy peak - these reads are a quarter of its wavefronts (same box: 828 -> 859 GS/s with the reference bands,
  // 750 -> 768 with 64 sub-channels, profiles/r02sl_ab_small_tmem.txt).  Two-pass plans whose team is a whole warp only:
  // tcgen05.ld / st are warp-wide (.sync.aligned), and where two half-warp teams share a warp (N = 256 / 512) the
  // unit epilogue lets them run different numbers of frames when K does not divide evenly - a diverged warp must not
  // issue them (measured there anyway with balanced K: +1.5 ... +2.5 %; the K = 7 / 10 launches hang).
  // -DCRN_SMALL_TMEM_MINN=<N>: smallest one-warp-per-frame size that does (A/B; 99999 = none).
#ifndef CRN_SMALL_TMEM_MINN
#define CRN_SMALL_TMEM_MINN 1024
#endif
  static constexpr bool TMEM_TW = (R2 == 1) && (T == 32) && (N >= CRN_SMALL_TMEM_MINN) && (NT % 128 == 0);
  static constexpr bool WIN_TMEM = TMEM_TW, ACC_TMEM = TMEM_TW;
  static constexpr int TM_TW = 0, TM_WIN = E, TM_ACC = 2 * E;   // first column of each table in a thread's TMEM row
  static constexpr int TMEM_COLS_PER_WARP = 3 * E;
  static constexpr int FIRST_RADIX = R0;                        // radix of the pass that consumes the window pairs
  static constexpr int TW_SMEM = TMEM_TW ? 0 : TW1 + TW2;       // float4 rows of the twiddle tables staged in shared memory
  // Window pairs from the table (false) or computed from two per-thread seeds (true; see HybridPlan::WIN_CALC).
  // -DCRN_WIN_CALC_SMALL=<smallest N that computes>: A/B switch for the one-warp-per-frame plans.
#ifdef CRN_WIN_CALC_SMALL
  static constexpr bool WIN_CALC = (N >= CRN_WIN_CALC_SMALL);
#else
  static constexpr bool WIN_CALC = false;
#endif
  // The window pairs (N/2 float2) are read through the L1 instead of shared memory at N >= 512: the streaming loads
  // are sensitive to how much of the SM's 256 KB is left as L1 (4 CTAs x 44 KB -> 196 KB carve-out; x 40 KB -> 164 KB:
  // +1.3 % at N = 1024, profiles/r02f_l1probe.txt; forcing the 228 KB carve-out costs 13-19 %).  -DCRN_PLAN_WIN_SMEM: A/B
#ifdef CRN_PLAN_WIN_SMEM
  static constexpr bool WIN_SMEM = !WIN_CALC && !WIN_TMEM;
#else
  static constexpr bool WIN_SMEM = (N < 512) && !WIN_CALC && !WIN_TMEM;
#endif
  // spectrum bin held in accumulator register m of team thread t after the last pass
  __host__ __device__ static constexpr int bin_of(int t, int m) { return t + T * m; }
  // where the epilogue parks bin b of a team's accumulators in shared memory (consecutive lanes -> consecutive bins here)
  __host__ __device__ static constexpr int pslot(int b) { return b; }
  // Bulk-copy (TMA) staging of the next frame pays off where a frame spans several warps and every
  // exchange is a multi-warp barrier (the hybrid plans); for the one-warp-per-frame sizes plain coalesced loads
  // are faster (same binary on the B200, packed-FP32 codelets: N = 1024 735.6 GS/s staged vs 751.7 plain,
  // N = 512 712 vs 770: the staged frame costs one more shared-memory read).
#ifdef CRN_SMALL_TMA  // A/B switch (build.py --variant): bulk-copy staging for the one-warp-per-frame plans too
  static constexpr bool TMA = true;
#else
  static constexpr bool TMA = (T > 64);
#endif
  // Software L2 prefetch one frame ahead.  Since the butterflies went to packed FP32 the one-warp-per-frame
  // kernels are no longer short of issue slots and hide the load latency themselves; the prefetch then only
  // adds L2 requests (measured: N = 512/1024 1-2 % faster without, N = 256 2 % faster with).
#ifdef CRN_SMALL_PREFETCH  // A/B switch: L2 prefetch one frame ahead at every size
  static constexpr bool PREFETCH = true;
#else
  static constexpr bool PREFETCH = (N < 512);
#endif
  static_assert(R0 * R1 * R2 == N, "radices must multiply to N");
  static_assert(E % R0 == 0 && E % R1 == 0 && E % R2 == 0, "E must be a multiple of every radix");
  static_assert(R0 >= 16, "first radix < 16 would bank-conflict the exchange");
  static_assert(T >= 16 && (T <= 32 ? 32 % T == 0 : T % 32 == 0), "team must tile a warp");
  static_assert(NT % UNIT_THREADS == 0 && (T <= 32 || UNITS <= 15), "units must tile the CTA (named barriers 1..15)");
  // epi_cta: the CTA-wide epilogue needs one row of segment sums and one feature row, not the per-unit two-slot ring
  static constexpr size_t smem_bytes(bool win, bool epi_cta, int seg_stride = CRN_MAX_SEGS, int band_stride = CRN_MAX_BANDS) {
    return sizeof(float4) * (size_t)TW_SMEM + sizeof(float2) * ((size_t)TEAMS * XSZ + (win && WIN_SMEM ? N / 2 : 0)) +
           sizeof(float) * (epi_cta ? seg_stride + band_stride : 2 * UNITS * seg_stride + band_stride * UNITS) +
           sizeof(int) * 4 * UNITS + 8 * TEAMS + 8;
  }
};

// Plan for N = C * 1024, C in {2, 4, 8}: "one cross-warp step, then every warp on its own".
//   pass A   radix-C decimation in frequency across the frame's C 1024-sample segments (registers
//            {i + r*G}, G = 32/C; window folded in): z_r[n], n < 1024;
//   exchange the only team-wide one: z_r goes to warp r;
//   pass B,C warp r runs the 1024-point 32x32 FFT of y_r[n] = z_r[n] W_N^(n r) with warp-local synchronisation only
//            and ends up holding bins C*k + r.  Both passes use twisted codelets (crn_fft_regs.cuh), and the DIF
//            twiddles W_N^(n r) cost nothing: with n = j + 32 q they split into (W_N^(32 r))^q - pass B's per-warp
//            w - and W_N^(j r), which multiplies pass C's input j and so merges with its W_1024^(j k):
//            w = W_N^(C k + r), one table column per team thread.
// Against the generic three-pass plans this halves the multi-warp barriers per frame, needs ~128 instead of
// 170-230 registers and a third less shared memory.
template <int N_, int TEAMS_, int MINB_>
struct HybridPlan {
  static constexpr int N = N_, E = 32, TEAMS = TEAMS_, MINB = MINB_;
  static constexpr int C = N / 1024;
  static constexpr int R0 = C, R1 = 32, R2 = 32;   // reported in the kernel name
  static constexpr int T = N / E;                  // 32*C threads = C warps per frame
  static constexpr int NT = T * TEAMS;
  static constexpr int PASSES = 3;
  static constexpr int PADSHIFT = 5;
  static constexpr int RS = 32 * (32 + 2);         // float2 slots of one warp's region (padded 32x32 exchange)
  static constexpr int XSZ = C * RS;
  // Tables in tensor memory (crn_fft_regs.cuh "TMEM as a per-thread table store").  ncu on the round-2 kernels: the
  // busiest pipe of the hybrid plans is the L1 / shared-memory data pipe (70 % at N = 8192, 84 % at 2048), and of its
  // ~430 wavefronts per 1024 samples 32 were window pairs, 32 pass-C twiddle rows, 8-12 pass-B rows - loop-invariant
  // values a thread re-reads every frame because 128 registers cannot hold them.  They now sit in the thread's own
  // TMEM columns (tcgen05.st once per CTA, tcgen05.ld per frame: a different pipe), the tables leave shared memory
  // (so the per-thread pass-C table of FOLD_C costs nothing at N = 8192 either and pass B loses its PRESCALE layer
  // there: -32 packed instructions per frame and thread), and 8192 x 64 sub-channels went 465 -> 565 GS/s together
  // with the other changes of this pass (profiles/r02s*_ab_*.txt).  -DCRN_NO_TMEM_TW: A/B.
#ifdef CRN_NO_TMEM_TW
  static constexpr bool TMEM_TW = false;
#else
  static constexpr bool TMEM_TW = true;
#endif
  // Window pairs in TMEM too (32 more columns per warp), in the order pass A consumes them.
  // -DCRN_WIN_TMEM_MAXC=<C>: largest C that reads the window from TMEM (A/B; 0 = never).
#ifndef CRN_WIN_TMEM_MAXC
#define CRN_WIN_TMEM_MAXC 8
#endif
  static constexpr bool WIN_TMEM = TMEM_TW && (C <= CRN_WIN_TMEM_MAXC);
  // The alternative that led there, kept as a switch: the Hann window computed instead of read.  Thread t needs
  // w[t + T m] = 1/2 - 1/2 cos(theta_t + m Delta), theta_t = 2 pi t / (N - 1), Delta = 2 pi T / (N - 1), and
  // cos(theta_t + m Delta) = cos(theta_t) cos(m Delta) - sin(theta_t) sin(m Delta) with cos / sin(m Delta) compile-time
  // immediates and (cos, sin)(theta_t) two per-thread registers: two scalar FFMA per window value (hann_calc).  Against
  // the shared-memory table: 8192 +6 %, 4096 +3 %, 2048 +2 % / -4 % (reference bands / all bins), one-warp plans 0;
  // against the TMEM table it loses 1-2 % (the FFMAs).  Used where WIN_TMEM is off and C >= CRN_WIN_CALC_MINC.
#ifndef CRN_WIN_CALC_MINC
#define CRN_WIN_CALC_MINC 4
#endif
  static constexpr bool WIN_CALC = (C >= CRN_WIN_CALC_MINC) && !WIN_TMEM;
  // Where the q-independent part W_N^(j r) of the DIF twiddles goes (see above).  FOLD_C: into pass C's w - free, but
  // the pass-C table then has one column per team thread (128 T bytes: 32 KB at N = 8192; in shared memory that pushed
  // the CTA past the 196 KB carve-out: measured -8 %, so round 2 used !FOLD_C there: pass B multiplies its column by
  // u = W_N^(j r) in its first stage, +32 packed instructions per frame and thread).  In TMEM the column is free.
#if defined(CRN_FOLD_C_ALL)   // A/B switches (build.py --variant)
  static constexpr bool FOLD_C = true;
#elif defined(CRN_FOLD_B_ALL)
  static constexpr bool FOLD_C = false;
#else
  static constexpr bool FOLD_C = (C < 8) || TMEM_TW;
#endif
  // C = 2: each warp keeps its own half of pass A's outputs (see the kernel); -DCRN_NO_OWN_SHARE: A/B switch
#ifdef CRN_NO_OWN_SHARE
  static constexpr bool OWN_SHARE = false;
#else
  static constexpr bool OWN_SHARE = (C == 2);
#endif
  // C = 4 / 8: a warp keeps the 1/C of pass A's outputs that is its own through a warp-uniform switch on the warp index
  // (one straight-line copy of the exchange per warp, static register indices in each).  -DCRN_KEEP_OWN=<max C>: A/B
#ifndef CRN_KEEP_OWN
#define CRN_KEEP_OWN 4
#endif
  static constexpr bool KEEP_OWN = (C > 2 && C <= CRN_KEEP_OWN);
  static constexpr int TW1 = FOLD_C ? 8 * T : 8 * 32;  // pass-C twisted-codelet table: 8 rows, one column per team thread / lane
  static constexpr int TW2 = 8 * C + (FOLD_C ? 0 : T);  // pass-B table: 8 rows, one column per warp (+ {u, u w^16} per thread)
  static constexpr int TW_SMEM = TMEM_TW ? 0 : TW1 + TW2;  // tables staged in shared memory (none when they live in TMEM)
  // The K-frame accumulators of the all-bins kernels in tensor memory as well (32 more columns): those kernels sit at
  // the 128-register cap with 32 accumulators + 64 data registers live through all three passes, and whether ptxas
  // spills, and how far it can hoist loads, moved N = 8192 by +-2 % from one harmless edit to the next.  With the
  // accumulators parked in TMEM between frames (read - add - write back in four 8-column batches behind the last pass)
  // the frame loop needs ~95 registers.  The kernels pruned to the reference band plan (10 accumulators) keep theirs in
  // registers.  -DCRN_NO_ACC_TMEM: A/B.
#ifdef CRN_NO_ACC_TMEM
  static constexpr bool ACC_TMEM = false;
#else
  static constexpr bool ACC_TMEM = TMEM_TW;
#endif
  // columns per warp: 8 rows x 4 for pass C, the same for pass B, 16 window pairs, 32 accumulators
  static constexpr int TMEM_COLS_PER_WARP = 64 + (WIN_TMEM || ACC_TMEM ? 32 : 0) + (ACC_TMEM ? 32 : 0);
  static constexpr int TM_TW = 0, TM_TWB = 32, TM_WIN = 64, TM_ACC = 96;  // first column of each table in a thread's TMEM row
  static constexpr int FIRST_RADIX = C;                                   // radix of the pass that consumes the window pairs
  static constexpr int UNIT_THREADS = T;
  static constexpr int UNITS = TEAMS;
  static constexpr int TEAMS_PER_UNIT = 1;
  static constexpr bool HYBRID = true;
  // Bulk-copy (TMA) staging of the team's next frame into its (by then idle) exchange regions: the copy
  // flies under the last pass and the accumulate, so the next frame's first pass starts from shared memory.
  // Measured in one binary with CRN_NO_TMA toggled, two frames per CTA and the window table in shared memory:
  // +4 % at N = 2048, +15 % at 4096, but -3 % at 8192, where the staged copy (one more shared-memory write and
  // read of the whole frame) meets a shared-memory pipe that is already ~70 % busy - there plain coalesced loads
  // behind the L2 prefetch win (435 -> 450 GS/s with 64 sub-channels, 467 -> 482 with the reference bands).
  // Second pass of round 2: once the prefetch no longer sits in front of the loads and the tables left shared
  // memory, plain loads win at every hybrid size (same box, CRN_NO_TMA toggled: 4096 613 -> 635 GS/s reference bands, 543 -> 557 all bins;
  // 2048 705 -> 709 / 618 -> 618): the staged frame's extra write + read of shared memory costs more than the load
  // latency it hides.  The staging code stays (parity-tested); -DCRN_TMA_MAXC=<C> compiles it in for C <= that (A/B).
#ifndef CRN_TMA_MAXC
#define CRN_TMA_MAXC 0
#endif
  static constexpr bool TMA = (C <= CRN_TMA_MAXC);
  // "My region is free" is an mbarrier arrival, not a team barrier.  The team exchange used to be bracketed by
  // two bar.sync: the first made sure every warp had gathered the previous frame's pass-B outputs out of its region
  // before anyone overwrites it - but a warp reaches that point right after its gather, a whole pass C + accumulate +
  // load + pass A before it needs the answer.  Each warp now arrives on the team's mbarrier after that gather and waits
  // for the phase just before its first store of the next frame (already complete by then), so pass A's stores need no
  // rendezvous; only the second barrier (stores visible -> gather) is left.  Shares mbars[] with the bulk-copy
  // staging, hence not both.  Measured (same box, profiles/r02sh_ab_tmem.txt): 2048 +1.9 % / +0.9 % (reference bands / all bins), 4096 +2.5 % / +1.4 %, 8192 -1.5 % /
  // -3.3 % (eight warps per team: without the first rendezvous the second one waits longer than both did), so C <= 4;
  // issuing pass A's stores codelet by codelet (EF_INTERLEAVE) changes nothing.  -DCRN_EARLY_FREE_MAXC=<C>: A/B.
#ifndef CRN_EARLY_FREE_MAXC
#define CRN_EARLY_FREE_MAXC 4
#endif
  static constexpr bool EARLY_FREE = !TMA && (C <= CRN_EARLY_FREE_MAXC);
#ifdef CRN_EF_INTERLEAVE  // A/B: pass A's stores issued codelet by codelet (generic exchange, C = 8)
  static constexpr bool EF_INTERLEAVE = true;
#else
  static constexpr bool EF_INTERLEAVE = false;
#endif
  // The window table (8 / 16 / 32 KB) lives in shared memory.  With one frame per CTA the two large ones were what
  // kept another CTA off the SM and were read through the read-only L1 path instead; with two frames per CTA
  // (4096: 2 CTAs/SM, 8192: 1) they fit, and shared memory is the faster home: 4096 +2 % (reference bands) /
  // +4 % (64 sub-channels), 8192 -1 % / +5 %.  -DCRN_WIN_L1 restores the L1 path from 4096 up (A/B).
#ifdef CRN_WIN_L1
  static constexpr bool WIN_SMEM = (N < 4096) && !WIN_CALC && !WIN_TMEM;
#else
  static constexpr bool WIN_SMEM = !WIN_CALC && !WIN_TMEM;
#endif
  static constexpr bool PREFETCH = true;           // hybrid plans: +2 % (2048) ... +5 % (8192) with
  static_assert(C == 2 || C == 4 || C == 8, "hybrid plans cover N = 2048, 4096, 8192");
  static_assert(UNITS <= 15, "named barriers 1..15");
  __host__ __device__ static constexpr int bin_of(int t, int m) { return C * ((t & 31) + 32 * m) + (t >> 5); }
  // Consecutive lanes hold bins C apart: parked at b they would hit 32 / C banks (ncu, N = 8192: 8 wavefronts per store,
  // 2 % of the kernel's shared-memory traffic); one pad word per 32 spreads them over all 32 banks.
  __host__ __device__ static constexpr int pslot(int b) { return b + (b >> 5); }
  static constexpr size_t smem_bytes(bool win, bool epi_cta, int seg_stride = CRN_MAX_SEGS, int band_stride = CRN_MAX_BANDS) {
    return sizeof(float4) * (size_t)TW_SMEM + sizeof(float2) * ((size_t)TEAMS * XSZ + (win && WIN_SMEM ? N / 2 : 0)) +
           sizeof(float) * (epi_cta ? seg_stride + band_stride : 2 * UNITS * seg_stride + band_stride * UNITS) +
           sizeof(int) * 4 * UNITS + 8 * TEAMS + 8;
  }
};

__device__ __forceinline__ float2 ld_stream(const float2 *p) {
  float2 v;
#ifdef CRN_LD_256
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#else
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
#endif
  return v;
}
// sc16 wire format: one 32-bit word = (I, Q) as int16.  The conversion is exact (|v| <= 32768 fits fp32);
// the 1/32768 scale is applied once to the finished features (folded into invK by the host).
__device__ __forceinline__ float2 unpack_sc16(unsigned v) {
  return make_float2((float)(short)(v & 0xffffu), (float)(short)(v >> 16));
}
__device__ __forceinline__ float2 ld_stream(const unsigned *p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return unpack_sc16(v);
}
__device__ __forceinline__ float2 ld_staged(const float2 *p) { return *p; }
__device__ __forceinline__ float2 ld_staged(const unsigned *p) { return unpack_sc16(*p); }

// Pull one frame (E*T*8 bytes, 128-byte lines) towards L2 ahead of use: lane t touches lines t + T*i.
// `line` points at this thread's first line (frame + 128 t bytes); `bytes_left` = frame bytes beyond it.
template <int E, int T, typename S>
__device__ __forceinline__ void prefetch_frame_l2(const S *line, int bytes_left) {
  constexpr int LINES_PER_THREAD = (E * (int)sizeof(S) + 127) / 128;  // E*T*sizeof(S)/128 lines over T threads
#pragma unroll
  for (int i = 0; i < LINES_PER_THREAD; i++) {
    if (128 * T * i < bytes_left)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(line) + 128 * T * i));
  }
}

// ---- TMA bulk-copy staging (cp.async.bulk + mbarrier): one elected thread per team pulls the team's next
// frame from HBM straight into the team's shared-memory buffer while the team is still computing ----------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_frame(void *dst_smem, const void *src_gmem, unsigned bytes,
                                               unsigned long long *bar) {
  // generic-proxy reads/writes of dst (ordered before this thread by the team barrier) must be complete
  // before the async proxy overwrites it
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, bytes);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMEM allocations are a power of two of columns, at least 32
__host__ __device__ constexpr unsigned tmem_alloc_cols(int need) {
  unsigned c = 32;
  while ((int)c < need) c *= 2;
  return c;
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

// Hann window value w[t + T M] of an N-point frame from the per-thread seeds (cos, sin)(2 pi t / (N - 1)):
// 1/2 - 1/2 cos(theta_t + M Delta), Delta = 2 pi T / (N - 1), by the angle-addition formula with cos / sin(M Delta)
// as immediates (liquid-dsp's symmetric Hann, CE extension; the table path evaluates the same formula in float on the
// host, the two agree to ~1e-7 absolute).
template <int N, int T, int M>
__device__ __forceinline__ float hann_calc(float2 seed) {
  constexpr float hc = (float)(-0.5 * cx_cos_turn(M * T, N - 1));
  constexpr float hs = (float)(0.5 * cx_sin_turn(M * T, N - 1));
  return fmaf(seed.y, hs, fmaf(seed.x, hc, 0.5f));
}
enum { WIN_NONE = 0, WIN_TABLE = 1, WIN_COMPUTED = 2, WIN_IN_TMEM = 3 };
// Position of window pair m0 (registers m0, m0 + E/2) in the thread's TMEM window columns: the R/2 pairs of codelet
// I = m0 % G are consecutive (pass A reads them with one tcgen05.ld), G = E / R.
__host__ __device__ constexpr int win_tmem_pos(int m0, int G, int R) { return (m0 % G) * (R / 2) + m0 / G; }


// Pass 0: E/R radix-R FFTs on registers {i + q*(E/R)}; the window (if any) rides on the first stage.
struct NoAfter {
  template <class I>
  __device__ __forceinline__ void operator()(I) const {}
};
// `after(I)` runs once codelet I's outputs are back in a[] (the hybrid plans store them to the exchange there).
template <int E, int R, int T, int WMODE, class After = NoAfter>
__device__ __forceinline__ void reg_pass_first(float2 (&a)[E], const float2 *__restrict__ winp, int t,
                                               float2 wseed = make_float2(0.f, 0.f), unsigned twin = 0,
                                               After after = After()) {
  constexpr int G = E / R;
  constexpr int LOG = ilog2(R);
  // WIN_IN_TMEM: the window pairs sit in TMEM in the order they are consumed (pair (I, Q) at position I R/2 + Q) and are
  // fetched PPC pairs per tcgen05.ld (8 columns; 4 for a radix-4 pass), the next batch in flight under the butterflies
  constexpr int PPC = (R == 4) ? 2 : 4;
  typename std::conditional<PPC == 4, TmemPending8, TmemPending4>::type wpend;
  float4 wq[2];
  if constexpr (WMODE == WIN_IN_TMEM) {
    static_assert(R >= 4 && ((E / 2) % PPC) == 0, "TMEM window: whole batches of pairs");
    if constexpr (PPC == 4) wpend = tmem_issue8(twin);
    else wpend = tmem_issue4(twin);
  }
  static_for<0, G>([&](auto I) {
    float2 v[R];
    static_for<0, R / 2>([&](auto Q) {
      constexpr int m0 = I.value + Q.value * G;  // partner is register m0 + E/2
      constexpr int br = bitrev(Q.value, LOG);   // even; bitrev(Q + R/2) == br + 1
      if constexpr (WMODE == WIN_IN_TMEM) {
        constexpr int p = I.value * (R / 2) + Q.value;  // position of this pair in the thread's window columns
        if constexpr (p % PPC == 0) {
          if constexpr (PPC == 4) {
            tmem_wait(wpend, wq[0], wq[1]);
            if constexpr (p + PPC < E / 2) wpend = tmem_issue8(twin + 2 * (p + PPC));
          } else {
            wq[0] = tmem_wait(wpend);
            if constexpr (p + PPC < E / 2) wpend = tmem_issue4(twin + 2 * (p + PPC));
          }
        }
        const float4 w4 = wq[(p % PPC) / 2];
        if constexpr (p & 1) butterfly_w_real(a[m0], a[m0 + E / 2], w4.z, w4.w, v[br], v[br + 1]);
        else butterfly_w_real(a[m0], a[m0 + E / 2], w4.x, w4.y, v[br], v[br + 1]);
      } else if constexpr (WMODE == WIN_TABLE) {
        const float2 w = winp[m0 * T + t];
        butterfly_w_real(a[m0], a[m0 + E / 2], w.x, w.y, v[br], v[br + 1]);
      } else if constexpr (WMODE == WIN_COMPUTED) {
        butterfly_w_real(a[m0], a[m0 + E / 2], hann_calc<E * T, T, m0>(wseed), hann_calc<E * T, T, m0 + E / 2>(wseed),
                         v[br], v[br + 1]);
      } else {
        v[br] = add2(a[m0], a[m0 + E / 2]);
        v[br + 1] = sub2(a[m0], a[m0 + E / 2]);
      }
    });
    fft_dit<R, 2>(v);
    static_for<0, R>([&](auto Q) { a[I.value + Q.value * G] = v[Q.value]; });
    after(I);
  });
}

// Pass p >= 1 of a Stockham plan (Ns = product of earlier radices): input q of column j = t + T*i carries the
// inter-pass twiddle W_{Ns*R}^(q (j mod Ns)) = w^q.  The codelet is "twisted" (crn_fft_regs.cuh): the twiddles ride
// on its butterflies, so the pass is R/2 log2(R) three-instruction butterflies and R/4 16-byte table reads per
// column - no separate twiddle layer.  `col0` is the table column of codelet 0; codelet i reads column
// col0[(T*i) & (NS-1)] (rows are RS float4 apart).
// PRESCALE: the whole column is also multiplied by a per-thread constant u (`usc` = {u, u w^(R/2)}); that costs two
// more packed instructions per first-stage butterfly (A = u a, p = A + (u tau) b, q = 2A - p).
template <int E, int R, int T, int NS, int RS, bool PRESCALE = false>
__device__ __forceinline__ void reg_pass_twisted(float2 (&a)[E], const float4 *__restrict__ twp, int t,
                                                 const float4 *__restrict__ usc = nullptr) {
  constexpr int G = E / R;
  constexpr int LOG = ilog2(R);
  static_for<0, G>([&](auto I) {
    const TwistedTable<RS> tw{twp + ((t + T * I.value) & (NS - 1))};
    float2 v[R];
    if constexpr (PRESCALE) {
      const float4 uu = *usc;
      static_for<0, R / 2>([&](auto Q) {
        constexpr int m0 = I.value + Q.value * G;
        constexpr int br = bitrev(Q.value, LOG);
        butterfly_w_cplx<false>(a[m0], a[m0 + E / 2], make_float2(uu.x, uu.y), make_float2(uu.z, uu.w), v[br], v[br + 1]);
      });
    } else {
    const float4 r0 = tw.template row<0>();
    const float2 tau = make_float2(r0.x, r0.y);  // w^(R/2): every first-stage butterfly
    static_for<0, R / 2>([&](auto Q) {
      constexpr int m0 = I.value + Q.value * G;  // partner is register m0 + E/2
      constexpr int br = bitrev(Q.value, LOG);   // even; bitrev(Q + R/2) == br + 1
      v[br] = a[m0];
      v[br + 1] = a[m0 + E / 2];
      butterfly_rt(v[br], v[br + 1], tau.x, -tau.y);
    });
    }
    fft_dit_twisted<R, 2>(v, tw);
    static_for<0, R>([&](auto Q) { a[I.value + Q.value * G] = v[Q.value]; });
  });
}

// The same pass with the thread's twiddle columns held in tensor memory.  E / R codelets per thread (codelet I on
// registers {I + q G}, its rows at taddr + I R): the one-warp-per-frame plans.
template <int E, int R>
__device__ __forceinline__ void reg_pass_twisted_tmem_multi(float2 (&a)[E], unsigned taddr) {
  constexpr int G = E / R;
  constexpr int LOG = ilog2(R);
  static_for<0, G>([&](auto I) {
    float2 v[R];
    static_for<0, R / 2>([&](auto Q) {
      constexpr int m0 = I.value + Q.value * G;
      constexpr int br = bitrev(Q.value, LOG);
      v[br] = a[m0];
      v[br + 1] = a[m0 + E / 2];
    });
    fft_dit_twisted_tmem<R>(v, taddr + R * I.value);
    static_for<0, R>([&](auto Q) { a[I.value + Q.value * G] = v[Q.value]; });
  });
}
// One radix-R codelet per thread (E == R): passes B and C of the hybrid plans.
template <int R>
__device__ __forceinline__ void reg_pass_twisted_tmem(float2 (&a)[R], unsigned taddr) {
  constexpr int LOG = ilog2(R);
  float2 v[R];
  static_for<0, R / 2>([&](auto Q) {
    constexpr int br = bitrev(Q.value, LOG);  // even; bitrev(Q + R/2) == br + 1
    v[br] = a[Q.value];
    v[br + 1] = a[Q.value + R / 2];
  });
#ifdef CRN_NO_TMEM_PIPE  // A/B: every row fetched and waited for where it is used (no row in flight under the butterflies)
  const TmemTwistedTable tw{taddr};
  const float4 r0 = tw.template row<0>();
  static_for<0, R / 2>([&](auto B) { butterfly_rt(v[2 * B.value], v[2 * B.value + 1], r0.x, -r0.y); });
  fft_dit_twisted<R, 2>(v, tw);
#else
  fft_dit_twisted_tmem<R>(v, taddr);
#endif
  static_for<0, R>([&](auto Q) { a[Q.value] = v[Q.value]; });
}

// Exchange rows are padded by two points: rows stay 16-byte aligned for the 16-byte stores of pass 0 and
// consecutive rows land in different banks (row stride 34 or 18 points -> conflict-free quarter-warps).
#define CRN_XPAD 2
template <int PADSHIFT>
__device__ __forceinline__ int xphys(int idx) {
  return idx + CRN_XPAD * (idx >> PADSHIFT);
}

// Scatter the outputs of a radix-R pass (Ns = product of earlier radices) into the exchange buffer,
// then gather this thread's points for the next pass.
template <int E, int R, int T, int NS, int PADSHIFT>
__device__ __forceinline__ void exchange(float2 (&a)[E], float2 *__restrict__ xb, int t, int team) {
  constexpr int G = E / R;
  team_sync<T>(team);  // previous readers of xb are done
  static_for<0, G>([&](auto I) {
    const int j = t + T * I.value;
    if constexpr (NS == 1 && R == (1 << PADSHIFT)) {
      // pass 0: this thread owns one padded row of R consecutive points -> 16-byte stores
      float4 *row = reinterpret_cast<float4 *>(xb + j * (R + CRN_XPAD));
      static_for<0, R / 2>([&](auto Q) {
        const float2 lo = a[I.value + (2 * Q.value) * G], hi = a[I.value + (2 * Q.value + 1) * G];
        row[Q.value] = make_float4(lo.x, lo.y, hi.x, hi.y);
      });
    } else {
      const int base = (j / NS) * (NS * R) + (j & (NS - 1));
      static_for<0, R>([&](auto Q) { xb[xphys<PADSHIFT>(base + Q.value * NS)] = a[I.value + Q.value * G]; });
    }
  });
  team_sync<T>(team);
  static_for<0, E>([&](auto M) { a[M.value] = xb[xphys<PADSHIFT>(t + T * M.value)]; });
}

template <int T, int UT>
__device__ __forceinline__ void unit_sync(int unit) {
  if constexpr (T > 32) asm volatile("bar.sync %0, %1;" ::"r"(unit + 1), "r"(UT) : "memory");
  else __syncwarp();
}

// MLP + first-match chain (or energy detector) for one decision, run by one warp; `fb` holds the features.
// The MLP (.cpp:200,214-235: double precision, logistic units, bias at index 0) is spread over lanes - hidden
// unit j on lane j, output k on lane k - so the warp pays for two dependent exp() instead of eight; every
// unit still sums in the reference's order, so the outputs are those of the serial loop bit for bit.
__device__ __forceinline__ void decide_and_store(const SenseParams &prm, const float *fb, long long g, int lane) {
  if (prm.decide == CRN_DECIDE_ANN) {
    const int j = (lane >= 1 && lane <= CRN_ANN_HIDDEN) ? lane : 1;
    double sum = prm.wih[0][j];
#pragma unroll
    for (int i = 1; i <= CRN_ANN_INPUTS; i++) sum += (double)fb[i - 1] * prm.wih[i][j];
    const double Hj = 1.0 / (1.0 + exp(-sum));  // Sigmoid_HA[j], valid on lanes 1..5
    const int k = (lane >= 1 && lane <= CRN_ANN_OUTPUTS) ? lane : 1;
    double so = prm.who[0][k];
#pragma unroll
    for (int jj = 1; jj <= CRN_ANN_HIDDEN; jj++) so += __shfl_sync(0xffffffffu, Hj, jj) * prm.who[jj][k];
    const double ok = 1.0 / (1.0 + exp(-so));    // Output[k], valid on lanes 1..3
    const double o1 = __shfl_sync(0xffffffffu, ok, 1), o2 = __shfl_sync(0xffffffffu, ok, 2),
                 o3 = __shfl_sync(0xffffffffu, ok, 3);
    if (lane == 0) {
      int dec = CRN_ALL_BUSY;  // .cpp:245-261
      if (o1 >= prm.threshold) dec = CRN_CH1_OCCUPIED;
      else if (o2 >= prm.threshold) dec = CRN_CH2_OCCUPIED;
      else if (o3 >= prm.threshold) dec = CRN_CH3_OCCUPIED;
      if (prm.ann) {
        prm.ann[3 * g + 0] = o1;
        prm.ann[3 * g + 1] = o2;
        prm.ann[3 * g + 2] = o3;
      }
      if (prm.decision) prm.decision[g] = dec;
      if (prm.mask) prm.mask[g] = dec ? (1ull << (dec - 1)) : 0ull;
    }
  } else {
    unsigned long long msk = 0ull;
    if (prm.decide == CRN_DECIDE_ENERGY) {
      float mn = 3.4e38f;
      for (int b = lane; b < prm.nbands; b += 32) mn = fminf(mn, fb[b]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      for (int b = lane; b < prm.nbands; b += 32)
        if ((double)fb[b] > prm.energy_factor * (double)mn) msk |= (1ull << b);
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

// Which accumulator registers a band plan needs.  After the last pass register m of a team thread holds a bin
// of the m-th N/E-wide slice of the spectrum (Plan::bin_of), so a band table that covers only part of the
// spectrum leaves whole registers unused - and with them the accumulate AND the butterflies of the last pass
// that feed nothing else (the compiler prunes them once the outputs are dead).  The reference engine's plan
// (CE_Predictive_Node.cpp:173-190: bins 0-15, 496-510, 55-84, 189-221, 300-309 of 512, scaled with N) touches
// 10 of 32 slices; kernels are instantiated for that mask and for "all registers" (any other table).
__host__ __device__ constexpr unsigned slices_of(int lo, int hi, int per_slice) {
  unsigned m = 0;
  for (int b = lo; b < hi; b++) m |= 1u << (b / per_slice);
  return m;
}
template <int E>
__host__ __device__ constexpr unsigned ref_acc_mask() {
  constexpr int W = 512 / E;  // slice width at N = 512
  return slices_of(0, 16, W) | slices_of(496, 511, W) | slices_of(55, 85, W) | slices_of(189, 222, W) |
         slices_of(300, 310, W);
}
template <int E>
__host__ __device__ constexpr unsigned full_acc_mask() { return E >= 32 ? 0xFFFFFFFFu : ((1u << E) - 1u); }

// EPI selects how the teams of a CTA share decision groups:
//   EPI_CTA  - the whole CTA works on one group (team q takes frames q, q+TEAMS, ...) and meets at three
//              CTA barriers per group to reduce; cheapest per group, best when a group is long (K >> TEAMS).
//   EPI_UNIT - `upg` units share a group and nobody waits: the last unit to arrive combines.  Keeps every
//              warp busy when groups are short (reference mode: K = 10) or K does not divide by TEAMS.
enum { EPI_CTA = 0, EPI_UNIT = 1 };

// Resident CTAs per SM an instantiation is compiled for (__launch_bounds__).  Plan::MINB is what every instantiation
// of a size fits; some fit more without spilling (tools/ptxas_report.py): the kernels pruned to the reference band plan
// lack ~22 accumulators (96 instead of 122-128 registers at N = 512 / 1024: one CTA more), the 16-point-per-thread plan of
// N = 256 runs in 80.  Taken only when that many CTAs also fit the SM's shared memory (228 KB, 1 KB reserved per CTA).
template <int N>
struct MoreCtas { static constexpr int cta_all = 0, cta_pruned = 0, unit_pruned = 0; };  // 0: Plan::MINB
#ifndef CRN_NO_MORE_CTAS  // A/B switch
template <> struct MoreCtas<256> { static constexpr int cta_all = 6, cta_pruned = 6, unit_pruned = 0; };
#ifdef CRN_MORE_CTAS_512_1024  // measured: 5 CTAs x 96 registers lose 12-18 % at N = 512 / 1024 (5 x 44 KB of shared
                               // memory push the carve-out to 228 KB and leave the streaming loads a 28 KB L1)
template <> struct MoreCtas<512> { static constexpr int cta_all = 0, cta_pruned = 5, unit_pruned = 5; };
template <> struct MoreCtas<1024> { static constexpr int cta_all = 0, cta_pruned = 5, unit_pruned = 0; };
#endif
#endif
template <class P, bool WIN, int EPI, unsigned AMASK>
__host__ __device__ constexpr int min_ctas() {
  constexpr bool pruned = AMASK != full_acc_mask<P::E>();
  constexpr int want = (EPI == EPI_CTA) ? (pruned ? MoreCtas<P::N>::cta_pruned : MoreCtas<P::N>::cta_all)
                                        : (pruned ? MoreCtas<P::N>::unit_pruned : 0);
  return (want > P::MINB && (P::smem_bytes(WIN, EPI == EPI_CTA) + 1024) * (size_t)want <= (size_t)228 * 1024) ? want : P::MINB;
}

template <class P, bool WIN, int DET, int EPI, bool SC16, unsigned AMASK>
__global__ void __launch_bounds__(P::NT, min_ctas<P, WIN, EPI, AMASK>()) sense_kernel(const SenseParams prm) {
  using sample_t = typename std::conditional<SC16, unsigned, float2>::type;  // one IQ sample in memory
  constexpr int SPL = 128 / (int)sizeof(sample_t);                           // samples per 128-byte line
  const sample_t *const iq = reinterpret_cast<const sample_t *>(prm.iq);
  constexpr int N = P::N, E = P::E, T = P::T, TEAMS = P::TEAMS, NT = P::NT;
  constexpr int UT = P::UNIT_THREADS, UNITS = P::UNITS, TPU = P::TEAMS_PER_UNIT;
  constexpr bool PREFETCH = P::PREFETCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *tw1 = reinterpret_cast<float4 *>(smem_raw);
  float4 *tw2 = tw1 + P::TW1;                                  // (both unused when the tables live in TMEM)
  float2 *xbuf = reinterpret_cast<float2 *>(tw1 + P::TW_SMEM);
  float2 *winp_s = xbuf + (size_t)TEAMS * P::XSZ;
  constexpr bool WSM = WIN && P::WIN_SMEM;
  const float2 *winp = WSM ? winp_s : prm.winp;  // window pairs: shared copy, or read-only global path
  constexpr bool ACC_IN_TMEM = P::ACC_TMEM && AMASK == full_acc_mask<P::E>() && P::E % 8 == 0;
  constexpr int WMODE = !WIN ? WIN_NONE : (P::WIN_TMEM ? WIN_IN_TMEM : (P::WIN_CALC ? WIN_COMPUTED : WIN_TABLE));
  // unit epilogue: [2][UNITS][seg_stride] partial sums + [UNITS][band_stride] features; CTA epilogue: one row each
  float *segpart = reinterpret_cast<float *>(winp_s + (WSM ? N / 2 : 0));
  const int SEGS = prm.seg_stride, BANDS = prm.band_stride;  // row lengths (the launch sized the allocation with them)
  float *featbuf = segpart + (EPI == EPI_CTA ? SEGS : 2 * UNITS * SEGS);
  int *cnt = reinterpret_cast<int *>(featbuf + (EPI == EPI_CTA ? BANDS : UNITS * BANDS));  // [2][UNITS] arrivals
  volatile int *done = cnt + 2 * UNITS;                                   // [2][UNITS] completed combines
  unsigned long long *mbars = reinterpret_cast<unsigned long long *>(
      (reinterpret_cast<uintptr_t>(cnt + 4 * UNITS) + 7) & ~(uintptr_t)7);  // [TEAMS] TMA arrival barriers

  const int tid = threadIdx.x;
  const int team = tid / T;
  const int t = tid % T;
  float2 wseed = make_float2(0.f, 0.f);  // (cos, sin)(2 pi t / (N - 1)): seeds of the computed window (hann_calc)
  if constexpr (WMODE == WIN_COMPUTED) {
    double sn, cs;
    sincospi(2.0 * (double)t / (double)(N - 1), &sn, &cs);
    wseed = make_float2((float)cs, (float)sn);
  }
  const int unit = tid / UT;            // reduction unit of this thread
  const int ut = tid % UT;              // thread index inside the unit
  const int upg = (EPI == EPI_CTA) ? UNITS : prm.upg;
  const int GL = (EPI == EPI_CTA) ? 1 : UNITS / upg;   // decision groups in flight per CTA
  const int gl = (EPI == EPI_CTA) ? 0 : unit / upg;    // which of them this unit works on
  const int fs = (EPI == EPI_CTA) ? team : (unit % upg) * TPU + (ut / T);  // frame slot inside the group
  const int FT = (EPI == EPI_CTA) ? TEAMS : upg * TPU;  // teams per group
  float2 *xb = xbuf + (size_t)team * P::XSZ;
  float *part = reinterpret_cast<float *>(xbuf + (size_t)(unit * TPU) * P::XSZ);  // unit's N floats

  // one-time table staging (persistent CTA: amortised over all its groups)
  for (int i = tid; i < P::TW_SMEM; i += NT) tw1[i] = prm.tw[i];
  unsigned tmem_tw = 0;  // TMEM address of this thread's twiddle rows (pass C: columns 0..31, pass B: 32..63)
  unsigned tacc = 0;     // ... of its accumulators (columns 96..127), ACC_IN_TMEM
  if constexpr (P::TMEM_TW) {
    // one warp allocates 64 columns per warp that shares a lane quarter (NT/128 of them; a power of two >= 32)
    constexpr unsigned COLS = tmem_alloc_cols(P::TMEM_COLS_PER_WARP * (NT / 128));
    static_assert(NT % 128 == 0 && COLS <= 512, "TMEM allocation shape");
    __shared__ unsigned tmem_base_s;
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned warp = (unsigned)tid >> 5;
    tmem_tw = tmem_base_s + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (unsigned)P::TMEM_COLS_PER_WARP;
    tacc = tmem_tw + P::TM_ACC;
    if constexpr (P::HYBRID) {
      // pass C (FOLD_C): column t of the first table; pass B: column r = t / 32 of the second (the same for a whole warp)
      static_for<0, 8>([&](auto R) { tmem_st4(tmem_tw + P::TM_TW + 4 * R.value, prm.tw[R.value * T + t]); });
      static_for<0, 8>([&](auto R) { tmem_st4(tmem_tw + 32 + 4 * R.value, prm.tw[P::TW1 + R.value * P::C + (t >> 5)]); });
    } else {
      // pass 1: codelet I of this thread works on column (t + T I) mod R0 of the table (R1/4 rows of R0 columns)
      static_for<0, E / P::R1>([&](auto I) {
        static_for<0, P::R1 / 4>([&](auto R) {
          tmem_st4(tmem_tw + P::TM_TW + P::R1 * I.value + 4 * R.value, prm.tw[R.value * P::R0 + ((t + T * I.value) & (P::R0 - 1))]);
        });
      });
    }
    if constexpr (WMODE == WIN_IN_TMEM) {
      // window pairs, two per 16-byte store, in the first pass's order (win_tmem_pos)
      constexpr int RF = P::FIRST_RADIX, G0 = E / RF;
      static_for<0, E / 4>([&](auto PP) {
        constexpr int p0 = 2 * PP.value, p1 = p0 + 1;
        constexpr int ma = (p0 % (RF / 2)) * G0 + p0 / (RF / 2), mb = (p1 % (RF / 2)) * G0 + p1 / (RF / 2);
        static_assert(win_tmem_pos(ma, G0, RF) == p0 && win_tmem_pos(mb, G0, RF) == p1, "window pair order");
        const float2 wa = prm.winp[ma * T + t], wb = prm.winp[mb * T + t];
        tmem_st4(tmem_tw + P::TM_WIN + 2 * p0, make_float4(wa.x, wa.y, wb.x, wb.y));
      });
    }
    tmem_wait_st();
  }
  if constexpr (WSM)
    for (int i = tid; i < N / 2; i += NT) winp_s[i] = prm.winp[i];
  for (int i = tid; i < 4 * UNITS; i += NT) cnt[i] = 0;
  const bool tma = P::TMA && (EPI == EPI_CTA) && prm.use_tma;
  static_assert(!(P::TMA && P::EARLY_FREE), "mbars[] serves either the bulk-copy staging or the region-free arrivals");
  if (tma && t == 0) {
    mbar_init(&mbars[team], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (P::EARLY_FREE) {
    if (t == 0) {
      mbar_init(&mbars[team], T / 32);  // one arrival per warp of the team
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  unsigned free_phase = 0;
  if constexpr (P::EARLY_FREE) {
    if ((tid & 31) == 0) mbar_arrive(&mbars[team]);  // phase 0: nothing to wait for before the first frame
  }

  const int L = prm.L, K = prm.K;
  const bool full = (L == N);
  // Work items (see "Group splitting" above).  The all-bins kernels sit exactly at the register cap, so the item
  // bookkeeping (split, frames per item, item count) stays in the parameter bank - operands, not registers - and
  // group / slice are derived from the item index only where they are needed.
  const int KP = (EPI == EPI_CTA) ? prm.kp : K;    // frames per work item
  const long long gstep = (long long)gridDim.x * GL;
  const unsigned frame_bytes = (unsigned)L * (unsigned)sizeof(sample_t);
  // first frame this team senses in work item w (group w / split, frames [slice*KP, (slice+1)*KP))
  auto first_frame = [&](long long w) -> const sample_t * {
    const int S = (EPI == EPI_CTA) ? prm.split : 1;
    const long long g = (S == 1) ? w : w / S;
    const int slice = (int)(w - g * S);
    return iq + ((size_t)g * (size_t)K + (size_t)slice * (size_t)KP + (size_t)fs) * (size_t)prm.stride;
  };
  unsigned tma_phase = 0;
  if (tma && t == 0 && fs < KP && (long long)blockIdx.x * GL + gl < prm.nwork)  // first item's first frame
    tma_load_frame(xb, first_frame((long long)blockIdx.x * GL + gl), frame_bytes, &mbars[team]);

  int it = 0;
  for (long long w = (long long)blockIdx.x * GL + gl; w < prm.nwork; w += gstep, it++) {
    float acc[E];  // (ACC_IN_TMEM: only the epilogue below sees them in registers)
#pragma unroll
    for (int m = 0; m < E; m++) acc[m] = 0.0f;
    if constexpr (ACC_IN_TMEM) {
      static_for<0, E / 8>([&](auto Cc) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        tmem_st8(tacc + 8 * Cc.value, z);
      });
      tmem_wait_st();
    }

    // frame pointers advance by plain 64-bit adds inside the loop; the multiplications happen once per item
    const size_t fstep = (size_t)FT * (size_t)prm.stride;  // samples between this team's frames
    const sample_t *x = first_frame(w) + t;
    for (int k = fs; k < KP; k += FT, x += fstep) {
      float2 a[E];
      if (tma) {
        // the frame was pulled into this team's buffer (linear layout) by the bulk copy issued one frame ago
        mbar_wait(&mbars[team], tma_phase);
        tma_phase ^= 1u;
        const sample_t *sx = reinterpret_cast<const sample_t *>(xb) + t;
        if (full) {
#pragma unroll
          for (int m = 0; m < E; m++) a[m] = ld_staged(sx + T * m);
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) a[m] = (t + T * m < L) ? ld_staged(sx + T * m) : make_float2(0.f, 0.f);
        }
      } else if (full) {
#pragma unroll
        for (int m = 0; m < E; m++) a[m] = ld_stream(x + T * m);
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) a[m] = (t + T * m < L) ? ld_stream(x + T * m) : make_float2(0.f, 0.f);
      }
      if constexpr (PREFETCH) {
        // Pull the frame this team senses next towards L2: k + FT of this item, else its first frame of the CTA's next
        // item.  One frame of compute covers the DRAM latency, so the loads above hit L2.  Issued AFTER this frame's
        // loads: the address used to be a select between two values kept alive across the whole frame loop -
        // spilled in the all-bins kernels, and the loads queued behind those local-memory reads (3 % of the N = 8192
        // kernel's warp time sat on that select); the item-boundary address is now rebuilt in its rare branch.
        // "+ (SPL-1) t" turns the per-thread sample pointer into a per-thread 128-byte line pointer.
        const sample_t *nx;
        if (k + FT < KP) nx = x + fstep + (SPL - 1) * t;
        else nx = (w + gstep < prm.nwork) ? first_frame(w + gstep) + t + (SPL - 1) * t : nullptr;
        if (nx) prefetch_frame_l2<E, T>(nx, (int)frame_bytes - 128 * t);
      }
      if constexpr (WMODE == WIN_COMPUTED) {
        // the 32 window values are loop-invariant; hoisted out of the frame loop they would be 32 more live registers
        // (spilled).  The empty asm makes the seeds opaque per frame, so the values are rebuilt where they are used.
        asm volatile("" : "+f"(wseed.x), "+f"(wseed.y));
      }
      if constexpr (P::EARLY_FREE) {
        // every warp of the team has gathered the previous frame's pass-B outputs: the regions may be overwritten
        mbar_wait(&mbars[team], free_phase);
        free_phase ^= 1u;
      }
      if constexpr (P::HYBRID) {
        constexpr int C = P::C, G = E / C;
        // pass A: radix-C across the frame's C segments (window folded in).  The DIF twiddles W_N^(n r) that turn
        // its outputs z_r[n] into the sub-sequences y_r[n] are not applied here: with n = j + 32 q they factor into
        // (W_N^(32 r))^q, which rides on pass B's butterflies, and W_N^(j r), which joins pass C's w.
        const int lane = t & 31;
        float2 *wb = xb + (t >> 5) * P::RS;
        if constexpr (P::OWN_SHARE) {
          // C = 2: a thread of warp w keeps the half of its outputs that belongs to its own warp (z_w) and hands
          // only the other half to the partner warp: half the exchange traffic (16 stores + 16 loads instead of 32
          // + 32).  Both warps run the same straight-line code:
          //  * the sign of the radix-2 butterfly is a per-thread constant (baked into the window pairs), so the
          //    "sum" register always holds z_w[n] and the "difference" register z_(1-w)[n];
          //  * warp 1 transforms its sub-sequence circularly shifted by 32 samples, y'_1[n'] = y_1[n' + 32] (|Y| is
          //    unchanged, the bins keep their places), so in BOTH warps the kept values are pass B's even inputs
          //    q = 2i and the received ones the odd inputs: warp 0 stores value i into slot i - 1 (mod 16) of warp
          //    1's region, negating the wrapped one (W_N^(n + 1024) = -W_N^n), warp 1 stores value i into slot i.
          const int w = t >> 5;
          const float sg = w ? -1.0f : 1.0f;
          float4 wq[2];  // WIN_IN_TMEM: four window pairs at a time, fetched one batch ahead
          TmemPending8 wpend;
          if constexpr (WMODE == WIN_IN_TMEM) wpend = tmem_issue8(tmem_tw + P::TM_WIN);
          static_for<0, E / 2>([&](auto I) {
            float2 p, q;
            if constexpr (WMODE == WIN_IN_TMEM) {
              if constexpr (I.value % 4 == 0) {
                tmem_wait(wpend, wq[0], wq[1]);
                if constexpr (I.value + 4 < E / 2) wpend = tmem_issue8(tmem_tw + P::TM_WIN + 2 * (I.value + 4));
              }
              const float4 w4 = wq[(I.value % 4) / 2];
              if constexpr (I.value & 1) butterfly_w_real(a[I.value], a[I.value + E / 2], w4.z, w4.w, p, q);
              else butterfly_w_real(a[I.value], a[I.value + E / 2], w4.x, w4.y, p, q);
            } else if constexpr (WMODE == WIN_TABLE) {
              const float2 wp = winp[I.value * T + t];  // { w[n], (+/-) w[n + N/2] }
              butterfly_w_real(a[I.value], a[I.value + E / 2], wp.x, wp.y, p, q);
            } else if constexpr (WMODE == WIN_COMPUTED) {
              butterfly_w_real(a[I.value], a[I.value + E / 2], hann_calc<N, T, I.value>(wseed),
                               sg * hann_calc<N, T, I.value + E / 2>(wseed), p, q);
            } else {
              p = fma2(a[I.value + E / 2], bc2(sg), a[I.value]);
              q = fma2(a[I.value + E / 2], bc2(-sg), a[I.value]);
            }
            a[I.value] = p;
            a[I.value + E / 2] = q;
          });
          a[E / 2] = mul2(a[E / 2], bc2(-sg));  // warp 0: the value that wraps around in warp 1's shifted sequence
          float2 *theirs = xb + (1 - w) * P::RS + lane;
          if constexpr (!P::EARLY_FREE) team_sync<T>(team);  // every warp of the team is done with the previous frame's regions
          theirs[w ? 0 : 32 * 15] = a[E / 2];
          static_for<1, E / 2>([&](auto I) { (theirs - (w ? 0 : 32))[32 * I.value] = a[I.value + E / 2]; });
          team_sync<T>(team);
          static_for<0, E / 2>([&](auto I) {  // descending: a[2i] = a[i] must not overwrite a value still needed
            constexpr int i = E / 2 - 1 - I.value;
            a[2 * i] = a[i];
          });
          static_for<0, E / 2>([&](auto I) { a[2 * I.value + 1] = wb[lane + 32 * I.value]; });
        } else if constexpr (P::KEEP_OWN) {
          // C = 4: warp w keeps z_w[n] for its own threads' n (8 of 32 values per thread) and sends the other three
          // sub-sequences: 24 stores + 24 loads per thread instead of 32 + 32.  As at C = 2, warp r transforms its
          // sub-sequence circularly shifted by 32 r samples (|Y| unchanged, bins in place), which makes the register
          // roles of pass B the same in every warp: kept values are inputs q = 4 i, the values from the warp d ahead
          // (mod 4) inputs 4 i + d.  The senders know everything statically inside a warp-uniform switch on their warp
          // index: the slot, and - for the one value per destination that wraps around the shifted sequence - the factor
          // W_N^(-1024 r) = j^r (a swap and a sign).  The kept values are picked out of the radix-4 outputs by selects.
          // (C = 8, -DCRN_KEEP_OWN=8: the wrap factor W_N^(-1024 r) is an eighth turn - one complex multiply by a
          // constant for the one wrapped value per destination.)
          reg_pass_first<E, C, T, WMODE>(a, winp, t, wseed, tmem_tw + P::TM_WIN);  // a[i + r G] = z_r[t + T i]
          const int w = t >> 5;
          if constexpr (!P::EARLY_FREE) team_sync<T>(team);  // every warp of the team is done with the previous frame's regions
          static_for<0, C>([&](auto W) {
            if (w == W.value) {
              static_for<0, C>([&](auto R) {
                if constexpr (R.value != W.value) {
                  constexpr int d = (W.value - R.value + C) % C;
                  float2 *dst = xb + R.value * P::RS + (d - 1) * G * 32 + lane;
                  if constexpr (W.value > R.value) {
                    static_for<0, G>([&](auto I) { dst[I.value * 32] = a[I.value + R.value * G]; });
                  } else {
                    const float2 v = a[R.value * G];  // n < 32 r: wraps to the end of warp r's shifted sequence
                    if constexpr (C == 4) {
                      dst[(G - 1) * 32] = R.value == 1 ? make_float2(-v.y, v.x)
                                                       : (R.value == 2 ? make_float2(-v.x, -v.y) : make_float2(v.y, -v.x));
                    } else {  // v * exp(+j 2 pi r / C)
                      constexpr float cr = (float)cx_cos_turn(R.value, C), sr = (float)cx_sin_turn(R.value, C);
                      dst[(G - 1) * 32] = make_float2(fmaf(v.x, cr, -v.y * sr), fmaf(v.x, sr, v.y * cr));
                    }
                    static_for<1, G>([&](auto I) { dst[(I.value - 1) * 32] = a[I.value + R.value * G]; });
                  }
                }
              });
            }
          });
          team_sync<T>(team);
          float2 own[G];
          static_for<0, G>([&](auto I) {
            float2 o = a[I.value];
            static_for<1, C>([&](auto R) {
              o.x = (w == R.value) ? a[I.value + R.value * G].x : o.x;
              o.y = (w == R.value) ? a[I.value + R.value * G].y : o.y;
            });
            own[I.value] = o;
          });
          static_for<0, G>([&](auto I) {
            a[C * I.value] = own[I.value];
            static_for<1, C>([&](auto D) { a[C * I.value + D.value] = wb[((D.value - 1) * G + I.value) * 32 + lane]; });
          });
        } else {
        // the one team-wide exchange: y_r[n] (n = t + T*i) goes to warp r's region, linear in n
        if constexpr (P::EARLY_FREE && P::EF_INTERLEAVE) {
          reg_pass_first<E, C, T, WMODE>(a, winp, t, wseed, tmem_tw + P::TM_WIN, [&](auto I) {
            static_for<0, C>([&](auto R) { xb[R.value * P::RS + t + T * I.value] = a[I.value + R.value * G]; });
          });
        } else {
          reg_pass_first<E, C, T, WMODE>(a, winp, t, wseed, tmem_tw + P::TM_WIN);
          if constexpr (!P::EARLY_FREE) team_sync<T>(team);  // every warp of the team is done with the previous frame's regions
          static_for<0, G>([&](auto I) {
            static_for<0, C>([&](auto R) { xb[R.value * P::RS + t + T * I.value] = a[I.value + R.value * G]; });
          });
        }
        team_sync<T>(team);
#pragma unroll
        for (int m = 0; m < E; m++) a[m] = wb[lane + 32 * m];
        }
        // warp r: 1024-point FFT of y_r, warp-local from here on (same code as the N = 1024 kernel)
        __syncwarp();  // the region is rewritten in padded layout below
        // pass B: column j = lane, input q carries (W_N^(32 r))^q - the same w for the whole warp (table column r)
        if constexpr (P::TMEM_TW) reg_pass_twisted_tmem<32>(a, tmem_tw + 32);
        else if constexpr (P::FOLD_C) reg_pass_twisted<E, 32, 32, 1, C>(a, tw2 + (t >> 5), 0);
        else reg_pass_twisted<E, 32, 32, 1, C, true>(a, tw2 + (t >> 5), 0, tw2 + 8 * C + t);
        exchange<E, 32, 32, 1, 5>(a, wb, lane, 0);
        if constexpr (P::EARLY_FREE) {
          __syncwarp();  // every lane's gather is done: this warp's region is free for the next frame's exchange
          if (lane == 0) mbar_arrive(&mbars[team]);
        }
        if (tma && k + FT < KP) {
          team_sync<T>(team);  // every warp of the team has gathered its points: the regions are idle
          if (t == 0) tma_load_frame(xb, x + fstep, frame_bytes, &mbars[team]);
        }
        // pass C: input q carries W_1024^(q lane) W_N^(q r) = (W_N^(C lane + r))^q: table column t = 32 r + lane
        if constexpr (P::TMEM_TW) reg_pass_twisted_tmem<32>(a, tmem_tw);
        else if constexpr (P::FOLD_C) reg_pass_twisted<E, 32, T, T, T>(a, tw1, t);
        else reg_pass_twisted<E, 32, 32, 32, 32>(a, tw1, lane);
      } else {
      // pass 0 (Ns = 1: no twiddles; window folded in)
      reg_pass_first<E, P::R0, T, WMODE>(a, winp, t, wseed, tmem_tw + P::TM_WIN);
      exchange<E, P::R0, T, 1, P::PADSHIFT>(a, xb, t, team);
      // after the frame's last exchange the buffer is free: start pulling this team's next frame now, so
      // the copy flies under the remaining butterflies and the accumulate
      auto stage_next = [&]() {
        if (tma && k + FT < KP) {
          team_sync<T>(team);  // every lane has gathered its points
          if (t == 0) tma_load_frame(xb, x + fstep, frame_bytes, &mbars[team]);  // (t == 0: x is the frame base)
        }
      };
      if constexpr (P::PASSES == 2) stage_next();
      // pass 1
      if constexpr (P::TMEM_TW) reg_pass_twisted_tmem_multi<E, P::R1>(a, tmem_tw + P::TM_TW);
      else reg_pass_twisted<E, P::R1, T, P::R0, P::R0>(a, tw1, t);
      if constexpr (P::PASSES == 3) {
        exchange<E, P::R1, T, P::R0, P::PADSHIFT>(a, xb, t, team);
        stage_next();
        reg_pass_twisted<E, P::R2, T, P::R0 * P::R1, P::R0 * P::R1>(a, tw2, t);
      }
      }  // !HYBRID
      // register m now holds bin P::bin_of(t, m)  (.cpp:152-154)
      if constexpr (ACC_IN_TMEM) {
        // accumulators live in tensor memory: eight at a time, the next batch in flight while this one is updated
        tmem_wait_st();  // last frame's write-back (long done) before these columns are read again
        TmemPending8 apend = tmem_issue8(tacc);
        static_for<0, E / 8>([&](auto Cc) {
          float v[8];
          tmem_wait(apend, v);
          if constexpr (Cc.value + 1 < E / 8) apend = tmem_issue8(tacc + 8 * (Cc.value + 1));
          static_for<0, 8>([&](auto J) {
            constexpr int m = 8 * Cc.value + J.value;
            if constexpr (DET == DET_MAGSQ) {
              v[J.value] = fmaf(a[m].x, a[m].x, v[J.value]);
              v[J.value] = fmaf(a[m].y, a[m].y, v[J.value]);
            } else {
              float s;
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(a[m].x, a[m].x, a[m].y * a[m].y)));
              v[J.value] += s;
            }
          });
          tmem_st8(tacc + 8 * Cc.value, v);
        });
      } else
      static_for<0, E>([&](auto M) {
        constexpr int m = M.value;
        if constexpr ((AMASK >> m) & 1u) {
          if constexpr (DET == DET_MAGSQ) {
            acc[m] = fmaf(a[m].x, a[m].x, acc[m]);
            acc[m] = fmaf(a[m].y, a[m].y, acc[m]);
          } else {
            float s;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(fmaf(a[m].x, a[m].x, a[m].y * a[m].y)));
            acc[m] += s;
          }
        }
      });
    }

    if constexpr (ACC_IN_TMEM) {
      tmem_wait_st();
      static_for<0, E / 8>([&](auto Cc) {
        float v[8];
        tmem_wait(tmem_issue8(tacc + 8 * Cc.value), v);
        static_for<0, 8>([&](auto J) { acc[8 * Cc.value + J.value] = v[J.value]; });
      });
    }
    const int S = (EPI == EPI_CTA) ? prm.split : 1;
    const long long g = (S == 1) ? w : w / S;  // decision group of this item
    const int ipart = (int)(w - g * S);        // which K/S-frame slice of it
    if constexpr (EPI == EPI_CTA) {
      // ---- per-group epilogue, CTA-wide ----------------------------------------------------------------
      float *segsum = segpart;  // [seg_stride]
      __syncthreads();          // every team finished reading its exchange buffer
      {
        float *mypart = reinterpret_cast<float *>(xb);  // N floats per team, aliasing the exchange buffer
        static_for<0, E>([&](auto M) {
          if constexpr ((AMASK >> M.value) & 1u) mypart[P::pslot(P::bin_of(t, M.value))] = acc[M.value];
        });
      }
      __syncthreads();
      {
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int NW = NT / 32;
        // four segments per trip so their loads and shuffle trees overlap instead of queueing up
        for (int sb = warp; sb < prm.nsegs; sb += 4 * NW) {
          float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int s = sb + u * NW;
            if (s < prm.nsegs)
              for (int i = prm.seg_lo[s] + lane; i < prm.seg_hi[s]; i += 32) {
#pragma unroll
                for (int q = 0; q < TEAMS; q++)
                  sum[u] += reinterpret_cast<const float *>(xbuf + (size_t)q * P::XSZ)[P::pslot(i)];
              }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 4; u++) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
          }
          if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 4; u++)
              if (sb + u * NW < prm.nsegs) segsum[sb + u * NW] = sum[u];
          }
        }
      }
      __syncthreads();
      if (tid < 32) {
        bool decide = true;
        if (S > 1) {
          // publish this item's segment sums; the item that arrives last adds the rows in part order
          float *row = prm.scratch + ((size_t)g * S + ipart) * (size_t)prm.nsegs;
          for (int sg = tid; sg < prm.nsegs; sg += 32) __stcg(row + sg, segsum[sg]);
          __threadfence();
          __syncwarp();
          int last = 0;
          if (tid == 0) last = (atomicAdd(prm.gcount + g, 1) == S - 1);
          decide = __shfl_sync(0xffffffffu, last, 0) != 0;
          if (decide) {
            __threadfence();
            const float *rows = prm.scratch + (size_t)g * S * (size_t)prm.nsegs;
            for (int sg = tid; sg < prm.nsegs; sg += 32) {
              float v = 0.0f;
              for (int q = 0; q < S; q++) v += __ldcg(rows + (size_t)q * prm.nsegs + sg);
              segsum[sg] = v;
            }
            if (tid == 0) prm.gcount[g] = 0;  // zero again for the next launch
            __syncwarp();
          }
        }
        if (decide) {
          for (int b = tid; b < prm.nbands; b += 32) {
            float m = 0.0f;
            if (prm.bands_contig) {
              for (int sg = prm.band_first[b]; sg < prm.band_first[b + 1]; sg++) m += segsum[sg];
            } else {
              for (int sg = 0; sg < prm.nsegs; sg++)
                if (prm.seg_band[sg] == b) m += segsum[sg];
            }
            m *= prm.invK;
            const float f = (prm.postop == CRN_POST_SQUARE_OF_SUM) ? m * m : m;  // .cpp:194-197
            featbuf[b] = f;
            prm.feat[(size_t)g * prm.nbands + b] = (prm.postop == CRN_POST_SUM_DB) ? 10.0f * log10f(f) : f;
          }
          __syncwarp();
          decide_and_store(prm, featbuf, g, tid);
        }
      }
      // the exchange buffers are free again (the last barrier above): pull the first frame of the next item
      if (tma && t == 0 && fs < KP && w + gstep < prm.nwork)
        tma_load_frame(xb, first_frame(w + gstep), frame_bytes, &mbars[team]);
      // nothing after the last barrier reads the exchange buffers, so the next group may start at once;
      // segsum/featbuf are rewritten only after the next group's barriers.
    } else {
    // ---- per-group epilogue -------------------------------------------------------------------------
    // (1) unit-local: accumulators -> shared (aliasing the unit's own exchange buffer) -> per-segment
    //     partial sums.  Only the unit's own threads synchronise.
    const int slot = it & 1;
    if constexpr (TPU == 2) {  // two half-warp teams hold the same bins: fold them first
      __syncwarp();
      static_for<0, E>([&](auto M) {
        if constexpr ((AMASK >> M.value) & 1u) acc[M.value] += __shfl_xor_sync(0xffffffffu, acc[M.value], 16);
      });
    }
    unit_sync<T, UT>(unit);  // every thread of the unit is done reading its exchange buffer
    if (ut < T) {
      static_for<0, E>([&](auto M) {
        if constexpr ((AMASK >> M.value) & 1u) part[P::pslot(P::bin_of(t, M.value))] = acc[M.value];
      });
    }
    unit_sync<T, UT>(unit);
    {
      const int wu = ut >> 5, lane = tid & 31;
      constexpr int NWU = UT / 32;
      // ring guard: the combine that last used this slot (two groups ago) must have finished reading
      if (lane == 0) {
        while (done[slot * UNITS + gl] != (it >> 1)) {
        }
      }
      __syncwarp();
      float *mine = segpart + ((size_t)slot * UNITS + unit) * SEGS;
      for (int sb = wu; sb < prm.nsegs; sb += 4 * NWU) {  // four segments per trip (independent chains)
        float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int s = sb + u * NWU;
          if (s < prm.nsegs)
            for (int i = prm.seg_lo[s] + lane; i < prm.seg_hi[s]; i += 32) sum[u] += part[P::pslot(i)];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < 4; u++) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (sb + u * NWU < prm.nsegs) mine[sb + u * NWU] = sum[u];
        }
      }
    }
    unit_sync<T, UT>(unit);  // partial sums written, `part` free again

    // (2) the unit's first warp announces arrival; the last unit of the group combines and decides
    if (ut < 32) {
      const int lane = ut;
      int last = 0;
      if (lane == 0) {
        __threadfence_block();
        last = (atomicAdd(&cnt[slot * UNITS + gl], 1) == upg - 1);
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence_block();
        float *fb = featbuf + unit * BANDS;
        const float *sp = segpart + ((size_t)slot * UNITS + (size_t)gl * upg) * SEGS;
        for (int b = lane; b < prm.nbands; b += 32) {
          float m = 0.0f;
          if (prm.bands_contig) {
            for (int s = prm.band_first[b]; s < prm.band_first[b + 1]; s++)
              for (int u = 0; u < upg; u++) m += sp[u * SEGS + s];
          } else {
            for (int s = 0; s < prm.nsegs; s++)
              if (prm.seg_band[s] == b)
                for (int u = 0; u < upg; u++) m += sp[u * SEGS + s];
          }
          m *= prm.invK;
          const float f = (prm.postop == CRN_POST_SQUARE_OF_SUM) ? m * m : m;  // .cpp:194-197
          fb[b] = f;
          prm.feat[(size_t)g * prm.nbands + b] = (prm.postop == CRN_POST_SUM_DB) ? 10.0f * log10f(f) : f;
        }
        __syncwarp();
        if (lane == 0) {
          cnt[slot * UNITS + gl] = 0;
          __threadfence_block();
          done[slot * UNITS + gl] = (it >> 1) + 1;  // partial sums consumed: the slot may be reused
        }
        decide_and_store(prm, fb, g, lane);
        __syncwarp();  // fb is reused by this warp's next combine
      }
    }
    }
  }
  if constexpr (P::TMEM_TW) {
    __syncthreads();  // every warp is done with its columns
    if (tid < 32) {
      constexpr unsigned COLS = tmem_alloc_cols(P::TMEM_COLS_PER_WARP * (NT / 128));
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_tw), "r"(COLS) : "memory");
    }
  }
}

}  // namespace crn
