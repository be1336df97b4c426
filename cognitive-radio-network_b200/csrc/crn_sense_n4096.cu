// sense_kernel instantiations for N = 4096 (radix 16 x 16 x 16, 16 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_4096(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<4096, 16, 16, 16, 16, 1, 2>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
