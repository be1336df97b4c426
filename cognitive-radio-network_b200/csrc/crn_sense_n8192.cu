// sense_kernel instantiations for N = 8192 (radix 32 x 16 x 16, 32 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_8192(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<8192, 32, 32, 16, 16, 1, 1>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
