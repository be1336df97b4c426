// sense_kernel instantiations for N = 1024 (radix 32 x 32 x 1, 32 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_1024(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<1024, 32, 32, 32, 1, 4, 4>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
