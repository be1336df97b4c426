// sense_kernel instantiations for N = 4096: hybrid plan (radix-4 across warps, then a 1024-point FFT per warp).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_4096(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
#ifdef CRN_T4096  // A/B switch (build.py --variant -DCRN_T4096=<teams per CTA> -DCRN_B4096=<CTAs per SM>)
  return launch_plan<HybridPlan<4096, CRN_T4096, CRN_B4096>>(prm, window, detector, grid, stream, geo);
#else
  return launch_plan<HybridPlan<4096, 2, 2>>(prm, window, detector, grid, stream, geo);
#endif
}
}  // namespace crn
