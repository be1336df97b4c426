// sense_kernel instantiations for N = 4096: hybrid plan (radix-4 across warps, then a 1024-point FFT per warp).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_4096(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<HybridPlan<4096, 1, 4>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
