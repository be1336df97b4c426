// Per-FFT-size launcher: picks the <window, detector> instantiation of sense_kernel for one Plan.
// Each crn_sense_n<N>.cu includes this once, so the six sizes compile in parallel.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "crn_internal.h"
#include "crn_sense_kernel.cuh"

namespace crn {

struct LaunchGeometry {
  int ctas_per_sm;   // resident CTAs per SM (occupancy query)
  int smem_bytes;
  int regs;
  int threads_per_cta, threads_per_frame, elems_per_thread, teams;
  int units, teams_per_unit;  // reduction units per CTA, teams per unit (see crn_sense_kernel.cuh)
  char name[64];
};

// signature shared by all sizes
typedef int (*sense_launch_fn)(const SenseParams &prm, int window, int detector, int grid,
                               cudaStream_t stream, LaunchGeometry *geo_only);

template <class P, bool WIN, int DET, int EPI, bool SC16, unsigned AMASK>
int launch_msk(const SenseParams &prm, int grid, cudaStream_t stream, LaunchGeometry *geo) {
  auto kern = sense_kernel<P, WIN, DET, EPI, SC16, AMASK>;
  // CRN_EXTRA_SMEM=<bytes>: development switch - unused dynamic shared memory on top, to move the SM's L1 / shared
  // carve-out without touching the kernel (the streaming loads are sensitive to the L1 size)
  static const size_t extra = getenv("CRN_EXTRA_SMEM") ? (size_t)atol(getenv("CRN_EXTRA_SMEM")) : 0;
  const size_t smem = P::smem_bytes(WIN, EPI == EPI_CTA, prm.seg_stride, prm.band_stride) + extra;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(CRN_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
  if (geo) {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return fail(CRN_ERR_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, P::NT, smem);
    if (e != cudaSuccess) return fail(CRN_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e));
    if constexpr (P::TMEM_TW) {
      // The occupancy API answers 1 CTA/SM for a kernel that allocates tensor memory, whatever it allocates; the
      // hardware keeps as many CTAs resident as registers, shared memory, threads and TMEM columns allow (ncu, N = 2048:
      // block limits registers 2 / shared memory 2, 15.6 of 16 warps active).  Count them here.
      int dev = 0, regs_sm = 0, smem_sm = 0, thr_sm = 0, smem_rsv = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
      cudaDeviceGetAttribute(&smem_rsv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
      const int regs_cta = ((fa.numRegs + 7) / 8) * 8 * P::NT;
      const size_t smem_cta = smem + fa.sharedSizeBytes + (size_t)smem_rsv;
      int n = regs_cta > 0 ? regs_sm / regs_cta : 1;
      if (smem_cta > 0 && (int)((size_t)smem_sm / smem_cta) < n) n = (int)((size_t)smem_sm / smem_cta);
      if (thr_sm / P::NT < n) n = thr_sm / P::NT;
      const int tmem = 512 / (int)tmem_alloc_cols(P::TMEM_COLS_PER_WARP * (P::NT / 128));
      if (tmem < n) n = tmem;
      if (n > occ) occ = n;
    }
    geo->ctas_per_sm = occ;
    geo->smem_bytes = (int)smem;
    geo->regs = fa.numRegs;
    geo->threads_per_cta = P::NT;
    geo->threads_per_frame = P::T;
    geo->elems_per_thread = P::E;
    geo->teams = P::TEAMS;
    geo->units = P::UNITS;
    geo->teams_per_unit = P::TEAMS_PER_UNIT;
    snprintf(geo->name, sizeof(geo->name), "sense_n%d_r%dx%dx%d_%s_%s_%s%s%s", P::N, P::R0, P::R1, P::R2,
             WIN ? "hann" : "rect", DET == DET_MAGSQ ? "magsq" : "mag", EPI == EPI_CTA ? "cta" : "unit",
             SC16 ? "_sc16" : "", AMASK == full_acc_mask<P::E>() ? "" : "_refbins");
    return CRN_OK;
  }
  kern<<<grid, P::NT, smem, stream>>>(prm);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRN_ERR_CUDA, "sense kernel launch: %s", cudaGetErrorString(e));
  return CRN_OK;
}

// Band tables that stay inside the reference engine's bin plan run the kernel pruned to those spectrum slices.
template <class P, bool WIN, int DET, int EPI, bool SC16>
int launch_fmt(const SenseParams &prm, int grid, cudaStream_t stream, LaunchGeometry *geo) {
  constexpr unsigned REF = ref_acc_mask<P::E>(), FULL = full_acc_mask<P::E>();
  return (prm.acc_mask & ~REF) == 0 ? launch_msk<P, WIN, DET, EPI, SC16, REF>(prm, grid, stream, geo)
                                    : launch_msk<P, WIN, DET, EPI, SC16, FULL>(prm, grid, stream, geo);
}

template <class P, bool WIN, int DET, int EPI>
int launch_epi(const SenseParams &prm, int grid, cudaStream_t stream, LaunchGeometry *geo) {
  return prm.sc16 ? launch_fmt<P, WIN, DET, EPI, true>(prm, grid, stream, geo)
                  : launch_fmt<P, WIN, DET, EPI, false>(prm, grid, stream, geo);
}

// prm.upg == 0 selects the CTA-wide epilogue, > 0 the unit epilogue with that many units per group.
template <class P, bool WIN, int DET>
int launch_one(const SenseParams &prm, int grid, cudaStream_t stream, LaunchGeometry *geo) {
  return prm.upg == 0 ? launch_epi<P, WIN, DET, EPI_CTA>(prm, grid, stream, geo)
                      : launch_epi<P, WIN, DET, EPI_UNIT>(prm, grid, stream, geo);
}

template <class P>
int launch_plan(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                LaunchGeometry *geo) {
  if (window == CRN_WINDOW_HANN) {
    return detector == CRN_DET_MAGSQ ? launch_one<P, true, DET_MAGSQ>(prm, grid, stream, geo)
                                     : launch_one<P, true, DET_MAG>(prm, grid, stream, geo);
  }
  return detector == CRN_DET_MAGSQ ? launch_one<P, false, DET_MAGSQ>(prm, grid, stream, geo)
                                   : launch_one<P, false, DET_MAG>(prm, grid, stream, geo);
}

// one per size, defined in crn_sense_n<N>.cu
int launch_sense_256(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);
int launch_sense_512(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);
int launch_sense_1024(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);
int launch_sense_2048(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);
int launch_sense_4096(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);
int launch_sense_8192(const SenseParams &, int, int, int, cudaStream_t, LaunchGeometry *);

// Radix plan per size (must match the Plan<> in crn_sense_n<N>.cu): the host builds the paired twiddle
// and window tables from it.
struct RadixPlan { int n, e, r0, r1, r2; bool hybrid; bool fold_c; bool own_share; };
inline RadixPlan radix_plan(int n) {
  switch (n) {
    case 256: return {256, 16, 16, 16, 1, false, false, false};
    case 512: return {512, 32, 32, 16, 1, false, false, false};
    case 1024: return {1024, 32, 32, 32, 1, false, false, false};
    case 2048: return {2048, 32, 2, 32, 32, true, HybridPlan<2048, 1, 1>::FOLD_C, HybridPlan<2048, 1, 1>::OWN_SHARE};
    case 4096: return {4096, 32, 4, 32, 32, true, HybridPlan<4096, 1, 1>::FOLD_C, HybridPlan<4096, 1, 1>::OWN_SHARE};
    case 8192: return {8192, 32, 8, 32, 32, true, HybridPlan<8192, 1, 1>::FOLD_C, HybridPlan<8192, 1, 1>::OWN_SHARE};
    default: return {0, 0, 0, 0, 0, false, false, false};
  }
}

}  // namespace crn
