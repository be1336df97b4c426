// sense_kernel instantiations for N = 2048 (radix 32 x 8 x 8, 32 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_2048(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<2048, 32, 32, 8, 8, 2, 3>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
