// sense_kernel instantiations for N = 2048: hybrid plan (radix-2 across warps, then a 1024-point FFT per warp).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_2048(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
#ifdef CRN_T2048  // A/B switch (build.py --variant -DCRN_T2048=<teams per CTA> -DCRN_B2048=<CTAs per SM>)
  return launch_plan<HybridPlan<2048, CRN_T2048, CRN_B2048>>(prm, window, detector, grid, stream, geo);
#else
  return launch_plan<HybridPlan<2048, 4, 2>>(prm, window, detector, grid, stream, geo);
#endif
}
}  // namespace crn
