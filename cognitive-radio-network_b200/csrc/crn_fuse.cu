// Hard-decision cooperative fusion of per-radio occupancy masks: OR / majority / AND across
// radios for every time slot.  The masks are what sense_kernel writes (CRN_DECIDE_ENERGY bit per band, or
// 1 << (channel-1) for the MLP decision), typically all-gathered from the GPUs that sensed the radios.
#include <cuda_runtime.h>

#include "crn_internal.h"

namespace {
__global__ void __launch_bounds__(256) fuse_kernel(const unsigned long long *__restrict__ masks, long long nradios,
                                                   long long nslots, int nbands, int mode,
                                                   unsigned long long *__restrict__ fused) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  // consecutive threads read consecutive slots of one radio row: coalesced
  unsigned long long any = 0ull, all = ~0ull;
  int votes[CRN_MAX_BANDS];
  if (mode == CRN_FUSE_MAJORITY)
    for (int c = 0; c < nbands; c++) votes[c] = 0;
  for (long long r = 0; r < nradios; r++) {
    const unsigned long long m = masks[r * nslots + s];
    any |= m;
    all &= m;
    if (mode == CRN_FUSE_MAJORITY)
      for (int c = 0; c < nbands; c++) votes[c] += (int)((m >> c) & 1ull);
  }
  unsigned long long out = (mode == CRN_FUSE_OR) ? any : all;
  if (mode == CRN_FUSE_MAJORITY) {
    out = 0ull;
    for (int c = 0; c < nbands; c++)
      if (2 * votes[c] > nradios) out |= (1ull << c);
  }
  const unsigned long long band_bits = nbands >= 64 ? ~0ull : ((1ull << nbands) - 1ull);
  fused[s] = out & band_bits;
}
}  // namespace

extern "C" int crn_fuse_masks_device(const uint64_t *d_masks, int64_t nradios, int64_t nslots, int32_t nbands,
                                     int32_t mode, uint64_t *d_fused, int32_t device, void *cuda_stream) {
  if (!d_masks || !d_fused || nradios < 1 || nslots < 0 || nbands < 1 || nbands > CRN_MAX_BANDS ||
      mode < CRN_FUSE_OR || mode > CRN_FUSE_AND)
    return crn::fail(CRN_ERR_INVALID, "crn_fuse_masks_device: bad argument");
  if (nslots == 0) return CRN_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  fuse_kernel<<<(unsigned)((nslots + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
      (const unsigned long long *)d_masks, nradios, nslots, nbands, mode, (unsigned long long *)d_fused);
  e = cudaGetLastError();
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "fuse kernel launch: %s", cudaGetErrorString(e));
  return CRN_OK;
}
