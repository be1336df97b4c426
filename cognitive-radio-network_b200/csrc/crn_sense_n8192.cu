// sense_kernel instantiations for N = 8192: hybrid plan (radix-8 across warps, then a 1024-point FFT per warp).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_8192(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<HybridPlan<8192, 1, 2>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
