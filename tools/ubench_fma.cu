// Issue-rate probe for the FP32 pipe of sm_100a: scalar FFMA (3-register and immediate form) against the
// packed FFMA2 / FADD2 forms the FFT codelets use.  Prints warp-instructions per cycle per SM sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_fma.cu -o tools/ubench_fma && tools/ubench_fma
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)), "l"(pk(c.x, c.y))); return upk(d); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y))); return upk(d); }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int CH = 8, IT = 512;
template <int V>
__global__ void __launch_bounds__(1024) probe(float2 *p, long long *cyc, int it) {
  float2 x[CH];
  for (int i = 0; i < CH; i++) x[i] = p[threadIdx.x + 32 * i];
  const float2 b = p[threadIdx.x + 1000], c = p[threadIdx.x + 2000];
  __syncthreads();
  long long t0 = clock64();
  for (int k = 0; k < it; k++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (V == 0) { x[i].x = fma1(x[i].x, b.x, c.x); x[i].y = fma1(x[i].y, b.y, c.y); }          // FFMA rrr (2 instr)
        if (V == 1) { x[i].x = fma1(x[i].x, 1.0001f, c.x); x[i].y = fma1(x[i].y, 0.9999f, c.y); }  // FFMA imm (2 instr)
        if (V == 2) x[i] = fma2(x[i], b, c);                                                       // FFMA2 rrr
        if (V == 3) x[i] = fma2(x[i], make_float2(1.0001f, 1.0001f), c);                           // FFMA2 imm
        if (V == 4) x[i] = fma2(make_float2(x[i].y, -x[i].x), make_float2(0.9999f, 0.9999f), c);   // FFMA2 swizzled imm
        if (V == 5) x[i] = add2(x[i], b);                                                          // FADD2
        if (V == 6) { x[i].x = x[i].x + b.x; x[i].y = x[i].y + b.y; }                               // FADD (2 instr)
      }
    }
  }
  long long t1 = clock64();
  float2 s = x[0];
  for (int i = 1; i < CH; i++) { s.x += x[i].x; s.y += x[i].y; }
  p[blockIdx.x * blockDim.x + threadIdx.x + 4096] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V>
void run(const char *name, int per_iter_instr, float2 *d, long long *dc, int threads) {
  int nsm = 148;
  probe<V><<<nsm, threads>>>(d, dc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<V><<<nsm, threads>>>(d, dc, IT);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
  double cy = 0; for (int i = 0; i < nsm; i++) cy += h[i]; cy /= nsm;
  double winstr_per_smsp = (double)IT * 4 * CH * per_iter_instr * (threads / 32) / 4.0;
  printf("%-22s warps/SMSP %2d: %.3f warp-instr/clk/SMSP  (%.3f flop-lanes/clk/SMSP, %.0f cycles, %.3f ms, err %d)\n", name, threads / 128,
         winstr_per_smsp / cy, winstr_per_smsp / cy * 32 * (V >= 2 && V <= 5 ? 2 : 1), cy, ms, (int)cudaGetLastError());
}
int main() {
  float2 *d; long long *dc;
  cudaMalloc(&d, (4096 + 148 * 1024) * sizeof(float2)); cudaMemset(d, 0, (4096 + 148 * 1024) * sizeof(float2));
  cudaMalloc(&dc, 148 * sizeof(long long));
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("FFMA rrr", 2, d, dc, threads);
    run<1>("FFMA imm", 2, d, dc, threads);
    run<2>("FFMA2 rrr", 1, d, dc, threads);
    run<3>("FFMA2 imm", 1, d, dc, threads);
    run<4>("FFMA2 swz imm", 1, d, dc, threads);
    run<5>("FADD2", 1, d, dc, threads);
    run<6>("FADD", 2, d, dc, threads);
  }
  return 0;
}
