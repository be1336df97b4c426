// sense_kernel instantiations for N = 8192: hybrid plan (radix-8 across warps, then a 1024-point FFT per warp).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_8192(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
#ifdef CRN_T8192  // A/B switch (build.py --variant -DCRN_T8192=<teams per CTA> -DCRN_B8192=<CTAs per SM>)
  return launch_plan<HybridPlan<8192, CRN_T8192, CRN_B8192>>(prm, window, detector, grid, stream, geo);
#else
  return launch_plan<HybridPlan<8192, 2, 1>>(prm, window, detector, grid, stream, geo);
#endif
}
}  // namespace crn
