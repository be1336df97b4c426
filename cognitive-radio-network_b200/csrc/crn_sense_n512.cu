// sense_kernel instantiations for N = 512 (radix 32 x 16 x 1, 32 points per thread).
#include "crn_launch.cuh"
namespace crn {
int launch_sense_512(const SenseParams &prm, int window, int detector, int grid, cudaStream_t stream,
                    LaunchGeometry *geo) {
  return launch_plan<Plan<512, 32, 32, 16, 1, 8, 4>>(prm, window, detector, grid, stream, geo);
}
}  // namespace crn
