// The occupancy predictor on its own: batched forward pass and on-device retraining of the reference's
// 4-5-3 logistic MLP (CE_Predictive_Node.hpp:62-73, forward pass .cpp:214-235, decision chain .cpp:245-261).
//
// forward  - one thread per decision, weights in the kernel-parameter bank, fp64 FMA + exp in the reference's
//            summation order (bias first, then inputs 1..4 / hidden units 1..5).  40 B in, 28 B out per decision:
//            HBM-bound, no tensor cores (35 MACs per decision).
// training - SURVEY 8f-4.  The reference ships only the result of its offline training (.cpp:74); this is plain
//            batch back-propagation for the same network.  One epoch = ann_grad_kernel (every thread walks
//            examples p, p + G*256, ...; 43 gradient sums + the error in fp64 registers; fixed-shape shuffle and
//            shared-memory tree -> one partial row per CTA) + ann_update_kernel (one CTA adds the partial rows in
//            order and applies momentum).  `check_every` epochs are captured once in a CUDA graph and replayed, so
//            the host is involved once per check, not twice per epoch.  Everything is deterministic for a given n.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "crn_internal.h"

namespace {

constexpr int NI = CRN_ANN_INPUTS, NH = CRN_ANN_HIDDEN, NO = CRN_ANN_OUTPUTS;
constexpr int NW_IH = (NI + 1) * NH;        // 25: wih[i][j], i = 0 (bias)..4, j = 1..5  -> i*NH + (j-1)
constexpr int NW = NW_IH + (NH + 1) * NO;   // 43: who[j][k], j = 0 (bias)..5, k = 1..3  -> NW_IH + j*NO + (k-1)
constexpr int NACC = NW + 1;                // gradient sums + the error
constexpr int TRAIN_THREADS = 256;

struct FlatWeights {
  double v[NW];
};

__host__ __device__ inline int ih(int i, int j) { return i * NH + (j - 1); }
__host__ __device__ inline int ho(int j, int k) { return NW_IH + j * NO + (k - 1); }

void flatten(const crn_ann_weights &w, double *v) {
  for (int i = 0; i <= NI; i++)
    for (int j = 1; j <= NH; j++) v[ih(i, j)] = w.wih[i][j];
  for (int j = 0; j <= NH; j++)
    for (int k = 1; k <= NO; k++) v[ho(j, k)] = w.who[j][k];
}
void unflatten(const double *v, crn_ann_weights *w) {
  memset(w, 0, sizeof(*w));  // row/column 0 of the "other" index stays unused, as in the reference
  for (int i = 0; i <= NI; i++)
    for (int j = 1; j <= NH; j++) w->wih[i][j] = v[ih(i, j)];
  for (int j = 0; j <= NH; j++)
    for (int k = 1; k <= NO; k++) w->who[j][k] = v[ho(j, k)];
}

// Hidden and output activations for one input vector (.cpp:214-235; same order of additions).
template <class W>
__device__ __forceinline__ void forward_one(const W &w, const double (&x)[NI + 1], double (&H)[NH + 1],
                                            double (&O)[NO + 1]) {
#pragma unroll
  for (int j = 1; j <= NH; j++) {
    double sum = w[ih(0, j)];
#pragma unroll
    for (int i = 1; i <= NI; i++) sum += x[i] * w[ih(i, j)];
    H[j] = 1.0 / (1.0 + exp(-sum));
  }
#pragma unroll
  for (int k = 1; k <= NO; k++) {
    double sum = w[ho(0, k)];
#pragma unroll
    for (int j = 1; j <= NH; j++) sum += H[j] * w[ho(j, k)];
    O[k] = 1.0 / (1.0 + exp(-sum));
  }
}

__global__ void __launch_bounds__(256) ann_forward_kernel(const FlatWeights w, double threshold,
                                                          const float *__restrict__ feat, long long n, int stride,
                                                          double *__restrict__ out, int *__restrict__ decision) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float *f = feat + p * stride;
  const double x[NI + 1] = {0.0, (double)f[0], (double)f[1], (double)f[2], (double)f[3]};  // .cpp:200
  double H[NH + 1], O[NO + 1];
  forward_one(w.v, x, H, O);
  if (out) {
    out[3 * p + 0] = O[1];
    out[3 * p + 1] = O[2];
    out[3 * p + 2] = O[3];
  }
  if (decision) {
    int dec = CRN_ALL_BUSY;  // .cpp:245-261
    if (O[1] >= threshold) dec = CRN_CH1_OCCUPIED;
    else if (O[2] >= threshold) dec = CRN_CH2_OCCUPIED;
    else if (O[3] >= threshold) dec = CRN_CH3_OCCUPIED;
    decision[p] = dec;
  }
}

struct TrainScale {
  double s[NI];
};

__global__ void __launch_bounds__(TRAIN_THREADS) ann_grad_kernel(const float *__restrict__ feat, int stride,
                                                                 const int *__restrict__ labels, long long n,
                                                                 const double *__restrict__ W, const TrainScale sc,
                                                                 double *__restrict__ partial) {
  __shared__ double sw[NW];
  __shared__ double red[TRAIN_THREADS / 32][NACC];
  for (int i = threadIdx.x; i < NW; i += blockDim.x) sw[i] = W[i];
  __syncthreads();
  double g[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) g[i] = 0.0;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += step) {
    const float *f = feat + p * stride;
    double x[NI + 1];
    x[0] = 1.0;
#pragma unroll
    for (int i = 1; i <= NI; i++) x[i] = (double)f[i - 1] * sc.s[i - 1];
    double H[NH + 1], O[NO + 1];
    forward_one(sw, x, H, O);
    H[0] = 1.0;
    const int label = labels[p];
    double dO[NO + 1];
#pragma unroll
    for (int k = 1; k <= NO; k++) {
      const double e = ((label == k) ? 1.0 : 0.0) - O[k];
      g[NW] += 0.5 * e * e;
      dO[k] = e * O[k] * (1.0 - O[k]);
    }
#pragma unroll
    for (int j = 0; j <= NH; j++) {
#pragma unroll
      for (int k = 1; k <= NO; k++) g[ho(j, k)] += H[j] * dO[k];
    }
#pragma unroll
    for (int j = 1; j <= NH; j++) {
      double sdow = 0.0;
#pragma unroll
      for (int k = 1; k <= NO; k++) sdow += sw[ho(j, k)] * dO[k];
      const double dH = sdow * H[j] * (1.0 - H[j]);
#pragma unroll
      for (int i = 0; i <= NI; i++) g[ih(i, j)] += x[i] * dH;
    }
  }
  // fixed-shape reduction: xor tree inside the warp, then the CTA's warps in order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    double v = g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double v = 0.0;
    for (int q = 0; q < TRAIN_THREADS / 32; q++) v += red[q][threadIdx.x];
    partial[(size_t)blockIdx.x * NACC + threadIdx.x] = v;
  }
}

// W += dW with dW = eta * (sum of partial gradients) / n + alpha * dW_previous; err[0] = E of this epoch.
__global__ void __launch_bounds__(64) ann_update_kernel(const double *__restrict__ partial, int nparts, double inv_n,
                                                        double eta, double alpha, double *__restrict__ W,
                                                        double *__restrict__ dW, double *__restrict__ err) {
  const int i = threadIdx.x;
  if (i >= NACC) return;
  double s = 0.0;
  for (int q = 0; q < nparts; q++) s += partial[(size_t)q * NACC + i];
  if (i < NW) {
    const double d = eta * s * inv_n + alpha * dW[i];
    dW[i] = d;
    W[i] += d;
  } else {
    err[0] = s;
  }
}

inline uint64_t mix64(uint64_t x) {  // splitmix64 step (the synthetic generator's hash, crn_synth.cu)
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

#define ANN_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      st = crn::fail(CRN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                    \
      goto done;                                                                                \
    }                                                                                           \
  } while (0)

}  // namespace

extern "C" int crn_ann_forward_device(const crn_ann_weights *w, double threshold, const float *d_feat, int64_t n,
                                      int32_t feat_stride, double *d_out, int32_t *d_decision, int32_t device,
                                      void *cuda_stream) {
  if (!w || !d_feat || n < 0 || feat_stride < NI || (!d_out && !d_decision))
    return crn::fail(CRN_ERR_INVALID, "crn_ann_forward_device: bad argument");
  if (n == 0) return CRN_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  FlatWeights fw;
  flatten(*w, fw.v);
  ann_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(
      fw, threshold, d_feat, n, feat_stride, d_out, d_decision);
  e = cudaGetLastError();
  if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "ann forward kernel launch: %s", cudaGetErrorString(e));
  return CRN_OK;
}

extern "C" int crn_ann_train_config_default(crn_ann_train_config *tc) {
  if (!tc) return crn::fail(CRN_ERR_INVALID, "crn_ann_train_config_default: null argument");
  memset(tc, 0, sizeof(*tc));
  tc->max_epochs = 2000;
  tc->check_every = 100;
  tc->eta = 0.5;
  tc->alpha = 0.9;
  tc->target_error = 0.0;
  for (int i = 0; i < NI; i++) tc->input_scale[i] = 1.0;
  tc->init_range = 0.5;
  tc->seed = 12;  // src/crts_cognitive_radio.cpp:754 seeds rand() with 12
  return CRN_OK;
}

extern "C" int crn_ann_train_device(const crn_ann_train_config *tc, const float *d_feat, int32_t feat_stride,
                                    const int32_t *d_labels, int64_t n, crn_ann_weights *w, double *final_error,
                                    int32_t *epochs_run, int32_t device, void *cuda_stream) {
  if (!tc || !d_feat || !d_labels || !w || n < 1 || feat_stride < NI || tc->max_epochs < 0 ||
      tc->check_every < 1 || !(tc->eta > 0.0) || tc->alpha < 0.0 || tc->alpha >= 1.0 || tc->init_range < 0.0)
    return crn::fail(CRN_ERR_INVALID, "crn_ann_train_device: bad argument");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));

  TrainScale sc;
  for (int i = 0; i < NI; i++) sc.s[i] = tc->input_scale[i] != 0.0 ? tc->input_scale[i] : 1.0;
  double hw[NW];
  if (tc->init_range > 0.0) {
    for (int i = 0; i < NW; i++) {
      const double u = (double)(mix64(tc->seed + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1)) >> 11) * (1.0 / 9007199254740992.0);
      hw[i] = (2.0 * u - 1.0) * tc->init_range;
    }
  } else {
    flatten(*w, hw);  // given for raw features: the network that sees scaled inputs has wih / scale
    for (int i = 1; i <= NI; i++)
      for (int j = 1; j <= NH; j++) hw[ih(i, j)] /= sc.s[i - 1];
  }

  int st = CRN_OK;
  int sms = 0;
  cudaStream_t stream = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  double *d_buf = nullptr;
  double err = 0.0;
  int epochs = 0;
  {
    ANN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int grid = (int)std::min<int64_t>((n + TRAIN_THREADS - 1) / TRAIN_THREADS, (int64_t)sms);
    // the examples may still be in flight on the caller's stream
    ANN_CUDA(cudaStreamSynchronize((cudaStream_t)cuda_stream));
    ANN_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    ANN_CUDA(cudaMalloc(&d_buf, sizeof(double) * (2 * NW + 1 + (size_t)grid * NACC)));
    double *d_W = d_buf, *d_dW = d_buf + NW, *d_err = d_buf + 2 * NW, *d_partial = d_buf + 2 * NW + 1;
    ANN_CUDA(cudaMemsetAsync(d_buf, 0, sizeof(double) * (2 * NW + 1), stream));
    ANN_CUDA(cudaMemcpyAsync(d_W, hw, sizeof(hw), cudaMemcpyHostToDevice, stream));
    const double inv_n = 1.0 / (double)n;
    auto epoch = [&](cudaStream_t s) {
      ann_grad_kernel<<<grid, TRAIN_THREADS, 0, s>>>(d_feat, feat_stride, d_labels, n, d_W, sc, d_partial);
      ann_update_kernel<<<1, 64, 0, s>>>(d_partial, grid, inv_n, tc->eta, tc->alpha, d_W, d_dW, d_err);
    };
    const int chunk = std::min(tc->check_every, std::max(tc->max_epochs, 1));
    if (tc->max_epochs >= chunk && tc->max_epochs > 0) {
      ANN_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < chunk; i++) epoch(stream);
      ANN_CUDA(cudaStreamEndCapture(stream, &graph));
      ANN_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    }
    while (epochs < tc->max_epochs) {
      const int left = tc->max_epochs - epochs;
      if (left >= chunk && exec) {
        ANN_CUDA(cudaGraphLaunch(exec, stream));
        epochs += chunk;
      } else {
        for (int i = 0; i < left; i++) epoch(stream);
        ANN_CUDA(cudaGetLastError());
        epochs += left;
      }
      ANN_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(double), cudaMemcpyDeviceToHost, stream));
      ANN_CUDA(cudaStreamSynchronize(stream));
      if (tc->target_error > 0.0 && err <= tc->target_error) break;
    }
    ANN_CUDA(cudaMemcpyAsync(hw, d_W, sizeof(hw), cudaMemcpyDeviceToHost, stream));
    ANN_CUDA(cudaStreamSynchronize(stream));
    for (int i = 1; i <= NI; i++)  // fold the input scale back: the engine feeds raw features (.cpp:200)
      for (int j = 1; j <= NH; j++) hw[ih(i, j)] *= sc.s[i - 1];
    unflatten(hw, w);
    if (final_error) *final_error = err;
    if (epochs_run) *epochs_run = epochs;
  }
done:
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  if (d_buf) cudaFree(d_buf);
  if (stream) cudaStreamDestroy(stream);
  return st;
}
