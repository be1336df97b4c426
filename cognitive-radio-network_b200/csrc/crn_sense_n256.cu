// sense_kernel instantiations for N = 256 (radix 16 x 16 x 1, 16 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_256(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<256, 16, 16, 16, 1, 8, 4>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
