// Read-bandwidth probe for the sensing kernel's access pattern (development tool, not part of the library).
// Each warp streams whole 8 KB frames (1024 complex samples); variants: 8-byte vs 16-byte loads per lane,
// with/without L2 prefetch one frame ahead, plus a configurable amount of dependent FMA work per frame.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/membw.cu -o /tmp/membw && /tmp/membw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int VEC, bool PF, int WORK>
__global__ void __launch_bounds__(128, 4) rd(const float2 *__restrict__ iq, long long nframes, float *out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 4;
  float acc = 0.f;
  for (long long f = warp; f < nframes; f += nwarps) {
    const float2 *x = iq + f * 1024;
    if (PF && f + nwarps < nframes) {
      const char *nx = (const char *)(iq + (f + nwarps) * 1024);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128 * lane));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128 * (lane + 32)));
    }
    float v[64];
    if (VEC == 8) {
#pragma unroll
      for (int m = 0; m < 32; m++)
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[2 * m]), "=f"(v[2 * m + 1]) : "l"(x + lane + 32 * m));
    } else {
#pragma unroll
      for (int m = 0; m < 16; m++)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[4 * m]), "=f"(v[4 * m + 1]), "=f"(v[4 * m + 2]), "=f"(v[4 * m + 3])
                     : "l"(x + 2 * lane + 64 * m));
    }
#pragma unroll
    for (int w = 0; w < WORK; w++)
#pragma unroll
      for (int i = 0; i < 64; i++) v[i] = fmaf(v[i], 1.0001f, v[(i + 1) & 63] * 0.5f);
#pragma unroll
    for (int i = 0; i < 64; i++) acc += v[i];
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int VEC, bool PF, int WORK>
void run(const float2 *d, long long nframes, float *out, const char *name) {
  for (int i = 0; i < 3; i++) rd<VEC, PF, WORK><<<592, 128>>>(d, nframes, out);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; i++) rd<VEC, PF, WORK><<<592, 128>>>(d, nframes, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("%-34s %.3f ms  %.1f GB/s  (%.1f GS/s)\n", name, ms, nframes * 8192.0 / ms / 1e6, nframes * 1024.0 / ms / 1e6);
}

int main() {
  const long long nframes = 976512;
  float2 *d;
  float *out;
  cudaMalloc(&d, nframes * 8192);
  cudaMalloc(&out, 4);
  cudaMemset(d, 0, nframes * 8192);
  run<8, false, 0>(d, nframes, out, "LDG.64  x32, no work");
  run<16, false, 0>(d, nframes, out, "LDG.128 x16, no work");
  run<8, true, 0>(d, nframes, out, "LDG.64  x32, L2 prefetch, no work");
  run<8, false, 4>(d, nframes, out, "LDG.64, 512 FMA+MUL/frame/thread");
  run<8, false, 8>(d, nframes, out, "LDG.64, 1024 FMA+MUL");
  run<8, true, 8>(d, nframes, out, "LDG.64, prefetch, 1024 FMA+MUL");
  run<16, false, 8>(d, nframes, out, "LDG.128, 1024 FMA+MUL");
  run<8, false, 6>(d, nframes, out, "LDG.64, 768 FMA+MUL");
  return 0;
}
