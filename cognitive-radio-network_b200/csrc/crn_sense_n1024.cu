// sense_kernel instantiations for N = 1024 (radix 32 x 32 x 1, 32 points per thread).
// CTA shape measured on the B200 (same box, 1e9 samples): 4 warps x 4 CTAs/SM 777 GS/s, 8 x 2 748, 2 x 8 623.
#include "crn_launch.cuh"
namespace crn {
int launch_sense_1024(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
#ifdef CRN_T1024  // A/B switch (build.py --variant -DCRN_T1024=<teams per CTA> -DCRN_B1024=<CTAs per SM>)
  return launch_plan<Plan<1024, 32, 32, 32, 1, CRN_T1024, CRN_B1024>>(prm, window, detector, grid, stream, geo);
#else
  return launch_plan<Plan<1024, 32, 32, 32, 1, 4, 4>>(prm, window, detector, grid, stream, geo);
#endif
}
}  // namespace crn
