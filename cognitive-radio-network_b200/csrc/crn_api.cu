// libcrnsense: handle, tables, pinned ring, batch paths.  See include/crnsense.h for the contract and
// the reference lines each entry point replaces.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <functional>
#include <vector>

#include "crn_internal.h"
#include "crn_launch.cuh"

#define CRN_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return crn::fail(CRN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                       __FILE__, __LINE__);                                                     \
  } while (0)

namespace {

// Results of one launch, device side (SoA) with a pinned host mirror.
struct ResultBuf {
  float *d_feat = nullptr;
  double *d_ann = nullptr;
  int32_t *d_dec = nullptr;
  unsigned long long *d_mask = nullptr;
  float *h_feat = nullptr;
  double *h_ann = nullptr;
  int32_t *h_dec = nullptr;
  unsigned long long *h_mask = nullptr;
  int64_t cap = 0;
};

// Scratch of a split launch (crn_sense_kernel.cuh "Group splitting"): per-part segment sums and per-group arrival
// counters (zero between launches).  Kernels that may overlap on the GPU must not share one: every ring slot,
// the batch-host pipeline and the batch-device entry point own theirs.
struct SplitBuf {
  float *d_scratch = nullptr;
  int *d_gcount = nullptr;
  size_t scratch_cap = 0, gcount_cap = 0;
};

struct RingSlot {
  unsigned char *h_iq = nullptr;  // pinned [K][L] samples in cfg.iq_format
  unsigned char *d_iq = nullptr;  // device mirror
  ResultBuf res;
  SplitBuf split;
  // where the decision of this slot is read from: the slot's own result buffer (index 0), or - for the members of a
  // crn_create_many pool - the pool slot's buffer at the member's index
  const ResultBuf *rv = nullptr;
  int64_t ri = 0;
  cudaEvent_t pool_done = nullptr;  // set while the slot's decision was launched by crn_submit_many (pool event)
  // device addresses of the pinned host buffers (zero-copy input / results), looked up once at create
  const float2 *z_iq = nullptr;
  float *z_feat = nullptr;
  double *z_ann = nullptr;
  int32_t *z_dec = nullptr;
  unsigned long long *z_mask = nullptr;
  cudaEvent_t done = nullptr;
  uint64_t first_frame = 0;
  int state = 0;  // 0 free/filling, 1 in flight
};

struct Pool;

}  // namespace

struct crn_handle {
  crn_config cfg;
  // crn_create_many: the handle is a member of a pool (shared pinned ring, tables, stream; see Pool below)
  Pool *pool = nullptr;
  int pool_index = -1;
  int device = 0;
  int num_sms = 0;
  int stride = 0;
  size_t sample_bytes = 8;  // 8 (CF32) or 4 (SC16)
  bool allow_tma = true;    // CRN_NO_TMA=1 (read once at create) forces plain loads: A/B measurements
  crn::sense_launch_fn launch = nullptr;
  crn::LaunchGeometry geo;
  crn::SenseParams base;  // everything but iq / outputs / ngroups
  float4 *d_tw = nullptr;
  float2 *d_win = nullptr;
  // group splitting (crn_sense_kernel.cuh): per-part segment sums and arrival counters, grown on demand
  int max_split = 16;       // CRN_SPLIT=<n> (read once at create) caps it; 1 disables splitting
  int force_split = 0;      // CRN_FORCE_SPLIT=<n>: development override (tuning of pick_split's cost model)
  double epilogue_frames = 2.0;  // pick_split: cost of one item's epilogue in units of one frame per team
  // Streaming path: from 2 MiB per decision the kernel reads the pinned slot over PCIe itself (transfer and FFTs
  // overlap frame by frame: 137 -> 127 us at 4 MiB); below that a host->device copy ahead of the kernel is quicker
  // (40 vs 43 us at 512 KiB, 27 vs 30 us at 40 KiB).  CRN_RING_COPY=1 / 0 forces the copy / the direct read (A/B).
  bool ring_zero_copy = false;
  SplitBuf host_split, dev_split;  // batch-host pipeline (h->stream) / batch-device launches (caller's stream)
  // Both are allocated at crn_create for the largest batch that is ever split (split_max_groups), so no launch
  // allocates or synchronises (crn_sense_batch_device stays asynchronous and legal under stream capture).  dev_split is
  // shared by the caller's streams: a split launch records dev_split_event, and a later split launch on a DIFFERENT
  // stream first waits for it, so two streams never share the arrival counters.
  int64_t split_max_groups = 0;
  cudaEvent_t dev_split_event = nullptr;
  cudaStream_t dev_split_stream = nullptr;
  bool dev_split_used = false;
  cudaStream_t stream = nullptr;     // streaming path + batch_host compute
  cudaStream_t copy_stream = nullptr;
  int64_t launches = 0;
  // streaming ring
  std::vector<RingSlot> ring;
  int fill_slot = 0, fill_frames = 0;
  int tail_slot = 0, inflight = 0;
  uint64_t frames_seen = 0;
  // batch-host staging (double buffered)
  int64_t chunk_groups = 0;
  unsigned char *h_stage[2] = {nullptr, nullptr};
  unsigned char *d_stage[2] = {nullptr, nullptr};
  ResultBuf stage_res[2];
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  cudaEvent_t stage_copied[2] = {nullptr, nullptr};
};

namespace {

// crn_create_many: n streaming handles of one configuration on one GPU that share everything a handle owns - tables,
// stream, and ONE pinned ring whose slot k holds the K frames of all n radios back to back ([radio][K][L], the layout
// of a batch of n decision groups).  A member handle is a view: its ring slots point into the pool's slots, its
// decisions are read from the pool's result arrays at its index.  crn_submit_many then senses the K-th-frame
// decisions of all members with one copy and ONE launch of the fused kernel over n groups.
struct Pool {
  crn_handle *proto = nullptr;  // an ordinary handle: kernel choice, tables, launch geometry, stream
  int n = 0;
  int live = 0;                 // members not yet destroyed
  size_t slot_bytes = 0;        // bytes of one member's K frames
  struct PSlot {
    unsigned char *h_iq = nullptr, *d_iq = nullptr;
    const float2 *z_iq = nullptr;
    ResultBuf res;              // n decisions, device arrays unused: the kernel writes the pinned mirrors
    float *z_feat = nullptr;
    double *z_ann = nullptr;
    int32_t *z_dec = nullptr;
    unsigned long long *z_mask = nullptr;
    SplitBuf split;             // n * max_split * nsegs: the pooled launch, or one member's own launch at its offset
    cudaEvent_t done = nullptr;
  };
  std::vector<PSlot> slots;
  std::vector<crn_handle *> members;
};

int alloc_results(ResultBuf &r, int64_t ngroups, int nbands) {
  r.cap = ngroups;
  CRN_CUDA(cudaMalloc(&r.d_feat, sizeof(float) * ngroups * nbands));
  CRN_CUDA(cudaMalloc(&r.d_ann, sizeof(double) * ngroups * 3));
  CRN_CUDA(cudaMalloc(&r.d_dec, sizeof(int32_t) * ngroups));
  CRN_CUDA(cudaMalloc(&r.d_mask, sizeof(unsigned long long) * ngroups));
  CRN_CUDA(cudaMallocHost(&r.h_feat, sizeof(float) * ngroups * nbands));
  CRN_CUDA(cudaMallocHost(&r.h_ann, sizeof(double) * ngroups * 3));
  CRN_CUDA(cudaMallocHost(&r.h_dec, sizeof(int32_t) * ngroups));
  CRN_CUDA(cudaMallocHost(&r.h_mask, sizeof(unsigned long long) * ngroups));
  return CRN_OK;
}
void free_results(ResultBuf &r) {
  cudaFree(r.d_feat);
  cudaFree(r.d_ann);
  cudaFree(r.d_dec);
  cudaFree(r.d_mask);
  cudaFreeHost(r.h_feat);
  cudaFreeHost(r.h_ann);
  cudaFreeHost(r.h_dec);
  cudaFreeHost(r.h_mask);
  r = ResultBuf();
}
int fetch_results_async(const ResultBuf &r, int64_t n, int nbands, cudaStream_t s) {
  CRN_CUDA(cudaMemcpyAsync(r.h_feat, r.d_feat, sizeof(float) * n * nbands, cudaMemcpyDeviceToHost, s));
  CRN_CUDA(cudaMemcpyAsync(r.h_ann, r.d_ann, sizeof(double) * n * 3, cudaMemcpyDeviceToHost, s));
  CRN_CUDA(cudaMemcpyAsync(r.h_dec, r.d_dec, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  CRN_CUDA(cudaMemcpyAsync(r.h_mask, r.d_mask, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, s));
  return CRN_OK;
}
void unpack_result(const ResultBuf &r, int64_t i, int nbands, uint64_t first_frame, crn_result *out) {
  memset(out, 0, sizeof(*out));
  out->first_frame = first_frame;
  out->decision = r.h_dec[i];
  out->nfeat = nbands;
  out->occupancy_mask = r.h_mask[i];
  for (int k = 0; k < 3; k++) out->ann_out[k] = r.h_ann[3 * i + k];
  for (int b = 0; b < nbands; b++) out->feat[b] = r.h_feat[i * nbands + b];
}

// Tables of the twisted codelets (crn_fft_regs.cuh), computed in double.  A radix-R codelet whose inputs carry
// w^q, w = exp(-j 2 pi phi), reads R/2 values  entry 0: w^(R/2);  entry m/4 + J: w^(R/m) W_m^J, J < m/4, m = 4..R,
// two per 16-byte row:  tw[(e/2) * ncols + col] = { entry e, entry e + 1 }  as (re, im, re, im).
void add_twisted_table(std::vector<float4> &tw, int r, int ncols, const std::function<double(int)> &phi_of_col) {
  auto entry = [&](int e, double phi) {
    int m = 2, j = 0;
    if (e > 0) {
      m = 4;
      while (m / 2 <= e) m *= 2;  // e in [m/4, m/2)
      j = e - m / 4;
    }
    const double turn = (double)(r / m) * phi + (double)j / (double)m;
    const double a = -2.0 * M_PI * (turn - floor(turn));
    return make_float2((float)cos(a), (float)sin(a));
  };
  for (int row = 0; row < r / 4; row++)
    for (int col = 0; col < ncols; col++) {
      const double phi = phi_of_col(col);
      const float2 a = entry(2 * row, phi), b = entry(2 * row + 1, phi);
      tw.push_back(make_float4(a.x, a.y, b.x, b.y));
    }
}
// Pass p of a Stockham plan (Ns = product of earlier radices, radix R): column j has w = W_{Ns R}^(j mod Ns).
void build_twiddles(const crn::RadixPlan &rp, std::vector<float4> &tw) {
  tw.clear();
  if (rp.hybrid) {
    // hybrid plan N = C * 1024 (crn_sense_kernel.cuh).  Pass C, FOLD_C: column t = 32 r + lane, w = W_N^(C lane + r);
    // otherwise column lane, w = W_1024^lane.  Then pass B, column r: w = W_N^(32 r); without FOLD_C followed by
    // { u, u w^16 } per team thread t = 32 r + lane, u = W_N^(lane r).
    const int c = rp.n / 1024;
    if (rp.fold_c)
      add_twisted_table(tw, 32, 32 * c, [&](int t) { return (double)(c * (t & 31) + (t >> 5)) / (double)rp.n; });
    else
      add_twisted_table(tw, 32, 32, [&](int j) { return (double)j / 1024.0; });
    add_twisted_table(tw, 32, c, [&](int r) { return (double)(32 * r) / (double)rp.n; });
    if (!rp.fold_c)
      for (int t = 0; t < 32 * c; t++) {
        const int lane = t & 31, r = t >> 5;
        const double au = -2.0 * M_PI * (double)(lane * r) / (double)rp.n;
        const double av = au - 2.0 * M_PI * (double)((16 * 32 * r) % rp.n) / (double)rp.n;
        tw.push_back(make_float4((float)cos(au), (float)sin(au), (float)cos(av), (float)sin(av)));
      }
    return;
  }
  add_twisted_table(tw, rp.r1, rp.r0, [&](int j) { return (double)j / (double)(rp.r0 * rp.r1); });
  if (rp.r2 > 1)
    add_twisted_table(tw, rp.r2, rp.r0 * rp.r1, [&](int j) { return (double)j / (double)(rp.r0 * rp.r1 * rp.r2); });
}

// Window in the kernel's register order: thread t of a team holds points t + T*m; the first butterfly
// stage pairs registers m and m + E/2:  winp[m*T + t] = { w[t + T*m], w[t + T*(m + E/2)] }, m < E/2.
// liquid-dsp hann(n, N) = 0.5 - 0.5 cos(2 pi n / (N - 1)), evaluated in float like liquid does.
void build_window_pairs(const crn::RadixPlan &rp, std::vector<float2> &wp) {
  const int N = rp.n, E = rp.e, T = N / E;
  std::vector<float> w(N);
  for (int n = 0; n < N; n++) w[n] = 0.5f - 0.5f * cosf((float)(2.0 * M_PI * (double)n) / (float)(N - 1));
  wp.resize(N / 2);
  // own-share hybrid plan (N = 2048): the threads of warp 1 (t >= 32) compute the radix-2 difference in the "sum"
  // register, i.e. their second weight carries a minus sign
  for (int m = 0; m < E / 2; m++)
    for (int t = 0; t < T; t++)
      wp[m * T + t] = make_float2(w[t + T * m], (rp.own_share && t >= 32 ? -1.0f : 1.0f) * w[t + T * (m + E / 2)]);
}

// Reduction units per decision group: the divisor of `units` that balances the K frames best over the
// group's teams (ties -> more units per group, i.e. fewer groups in flight per CTA and a shorter tail).
int pick_upg(int units, int teams_per_unit, int K) {
  int best = 1;
  double best_eff = -1.0;
  for (int d = 1; d <= units; d++) {
    if (units % d) continue;
    const int ft = d * teams_per_unit;
    const int rounds = (K + ft - 1) / ft;
    const double eff = (double)K / ((double)rounds * ft);
    if (eff >= best_eff - 1e-12) {
      best_eff = eff > best_eff ? eff : best_eff;
      best = d;
    }
  }
  return best;
}

// Grid of a launch.  Round 1 launched exactly as many CTAs as fit the GPU at once and let each stride over its share
// of the work items.  Equal shares are not equal times: the SMs of a B200 do not see HBM at the same speed (ncu: the
// per-SM active cycles of one launch spread by 6 %, and a resident warp slot is empty 17 % of the time), so the
// kernel ran at the pace of its slowest SM.  Large launches are therefore cut into `waves` times more CTAs than fit
// at once - the hardware block scheduler hands the next CTA to whichever SM frees a slot first, which is dynamic load
// balancing for the price of re-staging a few KB of tables per CTA.  Measured (profiles/r02i_gridmult.txt, same box):
// N = 1024 789 -> 836 GS/s (0.965 -> 1.02 of the measured copy peak), with every bin live 675 -> 742; 512 794 -> 834;
// reference mode 736 -> 782; 2048 638 -> 674; 4096 547 -> 578 at 8 waves; N = 8192 (one CTA per SM, compute/latency
// bound) gains nothing and stays persistent.  CRN_GRID_MULT=<waves> overrides (development).
int grid_waves(const crn_handle *h) {
  static const int forced = getenv("CRN_GRID_MULT") ? atoi(getenv("CRN_GRID_MULT")) : 0;
  if (forced >= 1) return forced;
  return h->cfg.nfft <= 2048 ? 16 : (h->cfg.nfft == 4096 ? 8 : 1);
}
int grid_for(const crn_handle *h, int64_t nwork) {
  int64_t g = (int64_t)h->num_sms * (h->geo.ctas_per_sm > 0 ? h->geo.ctas_per_sm : 1) * grid_waves(h);
  const int64_t gl = h->base.upg > 0 ? h->geo.units / h->base.upg : 1;  // groups a CTA works on at a time
  const int64_t need = (nwork + gl - 1) / gl;
  if (need < g) g = need;
  return (int)(g < 1 ? 1 : g);
}

// Work items per group for this launch (CTA epilogue only): the power of two <= max_split, with whole frames per
// team and item, that minimises rounds x (frames per team and item + epilogue), the epilogue counted as two
// frames.  Many groups -> 1 (nothing to gain); one decision on the streaming path -> as many items as K allows.
int pick_split(const crn_handle *h, int64_t ngroups) {
  if (h->base.upg != 0 || ngroups > h->split_max_groups) return 1;
  if (h->force_split > 0) return (h->cfg.navg % (h->force_split * h->geo.teams) == 0) ? h->force_split : 1;
  const int K = h->cfg.navg, teams = h->geo.teams;
  const int64_t ctas = (int64_t)h->num_sms * (h->geo.ctas_per_sm > 0 ? h->geo.ctas_per_sm : 1);
  int best = 1;
  double best_cost = 1e300;
  for (int sp = 1; sp <= h->max_split; sp *= 2) {
    if (K % (sp * teams)) break;
    const int64_t rounds = (ngroups * sp + ctas - 1) / ctas;
    const double cost = (double)rounds * ((double)(K / (sp * teams)) + h->epilogue_frames);
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = sp;
    }
  }
  return best;
}

// Scratch rows and arrival counters for a split launch.  Growing them waits for the device: earlier launches that
// own this buffer may still be using the old allocation.
int ensure_split_buffers(const crn_handle *h, SplitBuf &b, int64_t ngroups, int split) {
  const size_t need_s = (size_t)ngroups * split * h->cfg.nsegs, need_c = (size_t)ngroups;
  if (need_s <= b.scratch_cap && need_c <= b.gcount_cap) return CRN_OK;
  CRN_CUDA(cudaDeviceSynchronize());
  if (need_s > b.scratch_cap) {
    cudaFree(b.d_scratch);
    b.d_scratch = nullptr;
    b.scratch_cap = 0;
    CRN_CUDA(cudaMalloc(&b.d_scratch, sizeof(float) * need_s));
    b.scratch_cap = need_s;
  }
  if (need_c > b.gcount_cap) {
    cudaFree(b.d_gcount);
    b.d_gcount = nullptr;
    b.gcount_cap = 0;
    CRN_CUDA(cudaMalloc(&b.d_gcount, sizeof(int) * need_c));
    CRN_CUDA(cudaMemset(b.d_gcount, 0, sizeof(int) * need_c));
    b.gcount_cap = need_c;
  }
  return CRN_OK;
}
void free_staging(crn_handle *h) {
  for (int i = 0; i < 2; i++) {
    cudaFreeHost(h->h_stage[i]);
    cudaFree(h->d_stage[i]);
    h->h_stage[i] = h->d_stage[i] = nullptr;
    free_results(h->stage_res[i]);
    if (h->stage_done[i]) cudaEventDestroy(h->stage_done[i]);
    if (h->stage_copied[i]) cudaEventDestroy(h->stage_copied[i]);
    h->stage_done[i] = h->stage_copied[i] = nullptr;
  }
  h->chunk_groups = 0;
}
void free_split_buffers(SplitBuf &b) {
  cudaFree(b.d_scratch);
  cudaFree(b.d_gcount);
  b = SplitBuf();
}

// Fills the launch-shape dependent fields of p (split, scratch) and returns the grid size through *grid.
int shape_launch(crn_handle *h, crn::SenseParams &p, int64_t ngroups, SplitBuf &buf, int *grid) {
  p.split = pick_split(h, ngroups);
  p.kp = h->cfg.navg / p.split;
  p.nwork = ngroups * p.split;
  p.scratch = nullptr;
  p.gcount = nullptr;
  if (p.split > 1) {
    // preallocated at crn_create (pick_split never splits a batch larger than split_max_groups)
    if ((size_t)ngroups * p.split * h->cfg.nsegs > buf.scratch_cap || (size_t)ngroups > buf.gcount_cap)
      return crn::fail(CRN_ERR_INVALID, "split scratch too small for %lld groups x %d", (long long)ngroups, p.split);
    p.scratch = buf.d_scratch;
    p.gcount = buf.d_gcount;
  }
  *grid = grid_for(h, ngroups * p.split);
  return CRN_OK;
}

int launch(crn_handle *h, SplitBuf &split, const void *d_iq, int64_t ngroups, float *d_feat, double *d_ann,
           int32_t *d_dec, unsigned long long *d_mask, cudaStream_t s) {
  if (ngroups <= 0) return CRN_OK;
  crn::SenseParams p = h->base;
  p.iq = static_cast<const float2 *>(d_iq);
  p.feat = d_feat;
  p.ann = d_ann;
  p.decision = d_dec;
  p.mask = d_mask;
  p.ngroups = ngroups;
  // bulk-copy (TMA) staging needs 16-byte aligned frame addresses and sizes; anything else uses plain loads
  p.use_tma = (p.upg == 0) && h->allow_tma && ((reinterpret_cast<uintptr_t>(d_iq) & 15) == 0) &&
              ((h->stride * h->sample_bytes) % 16 == 0) && ((h->cfg.frame_len * h->sample_bytes) % 16 == 0);
  int grid = 1;
  int st = shape_launch(h, p, ngroups, split, &grid);
  if (st != CRN_OK) return st;
  const bool shared_scratch = (p.split > 1) && (&split == &h->dev_split);
  if (shared_scratch && h->dev_split_used && h->dev_split_stream != s)
    CRN_CUDA(cudaStreamWaitEvent(s, h->dev_split_event, 0));  // the other stream's split launch owns the counters
  st = h->launch(p, h->cfg.window, h->cfg.detector, grid, s, nullptr);
  if (st == CRN_OK) h->launches++;
  if (st == CRN_OK && shared_scratch) {
    CRN_CUDA(cudaEventRecord(h->dev_split_event, s));
    h->dev_split_stream = s;
    h->dev_split_used = true;
  }
  return st;
}

}  // namespace

extern "C" {

int crn_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return crn::fail(CRN_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int crn_create(const crn_config *cfg, crn_handle **out) {
  if (!out) return crn::fail(CRN_ERR_INVALID, "crn_create: null out pointer");
  *out = nullptr;
  int st = crn_config_validate(cfg);
  if (st != CRN_OK) return st;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return crn::fail(CRN_ERR_NO_DEVICE, "no CUDA device (%s); libcrnsense has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
  if (cfg->device < 0 || cfg->device >= ndev)
    return crn::fail(CRN_ERR_NO_DEVICE, "device %d out of range (have %d)", cfg->device, ndev);
  CRN_CUDA(cudaSetDevice(cfg->device));

  crn_handle *h = new (std::nothrow) crn_handle();
  if (!h) return crn::fail(CRN_ERR_NOMEM, "out of host memory");
  // any early return below (CRN_CUDA) releases what was allocated so far
  struct Guard {
    crn_handle *h;
    ~Guard() { if (h) crn_destroy(h); }
  } guard{h};
  h->cfg = *cfg;
  h->device = cfg->device;
  h->stride = cfg->frame_stride > 0 ? cfg->frame_stride : cfg->frame_len;
  h->sample_bytes = cfg->iq_format == CRN_IQ_SC16 ? 4 : 8;
  if (h->cfg.ring_slots == 0) h->cfg.ring_slots = 4;
  cudaDeviceProp prop;
  CRN_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  h->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    return crn::fail(CRN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                     cfg->device, prop.major, prop.minor);
  }

  switch (cfg->nfft) {
    case 256: h->launch = crn::launch_sense_256; break;
    case 512: h->launch = crn::launch_sense_512; break;
    case 1024: h->launch = crn::launch_sense_1024; break;
    case 2048: h->launch = crn::launch_sense_2048; break;
    case 4096: h->launch = crn::launch_sense_4096; break;
    case 8192: h->launch = crn::launch_sense_8192; break;
    default: return crn::fail(CRN_ERR_UNSUPPORTED, "nfft %d not compiled", cfg->nfft);
  }

  // tables
  std::vector<float4> tw;
  build_twiddles(crn::radix_plan(cfg->nfft), tw);
  CRN_CUDA(cudaMalloc(&h->d_tw, sizeof(float4) * tw.size()));
  CRN_CUDA(cudaMemcpy(h->d_tw, tw.data(), sizeof(float4) * tw.size(), cudaMemcpyHostToDevice));
  if (cfg->window == CRN_WINDOW_HANN) {
    std::vector<float2> wp;
    build_window_pairs(crn::radix_plan(cfg->nfft), wp);
    CRN_CUDA(cudaMalloc(&h->d_win, sizeof(float2) * wp.size()));
    CRN_CUDA(cudaMemcpy(h->d_win, wp.data(), sizeof(float2) * wp.size(), cudaMemcpyHostToDevice));
  }

  crn::SenseParams &b = h->base;
  memset(&b, 0, sizeof(b));
  b.tw = h->d_tw;
  b.winp = h->d_win;
  b.upg = 1;
  b.L = cfg->frame_len;
  b.stride = h->stride;
  b.K = cfg->navg;
  // sc16 samples are converted to float unscaled (exact); their 1/32768 (or its square for |X|^2) rides on 1/K
  {
    double scale = 1.0 / (double)cfg->navg;
    if (cfg->iq_format == CRN_IQ_SC16) scale *= (cfg->detector == CRN_DET_MAGSQ) ? (1.0 / 32768.0) * (1.0 / 32768.0) : (1.0 / 32768.0);
    b.invK = (float)scale;
    b.sc16 = cfg->iq_format == CRN_IQ_SC16;
  }
  b.nbands = cfg->nbands;
  b.nsegs = cfg->nsegs;
  b.seg_stride = (cfg->nsegs + 7) & ~7;    // shared-memory scratch rows of the epilogue are sized to the band plan
  b.band_stride = (cfg->nbands + 7) & ~7;
  b.postop = cfg->postop;
  b.decide = cfg->decide;
  b.threshold = cfg->ann_threshold;
  b.energy_factor = cfg->energy_factor;
  memcpy(b.wih, cfg->ann_wih, sizeof(b.wih));
  memcpy(b.who, cfg->ann_who, sizeof(b.who));
  for (int s = 0; s < cfg->nsegs; s++) {
    b.seg_band[s] = (short)cfg->segs[s].band;
    b.seg_lo[s] = (short)cfg->segs[s].lo;
    b.seg_hi[s] = (short)cfg->segs[s].hi;
  }
  {  // segments listed band by band (non-decreasing band index)?  then band b owns a contiguous run of entries
    b.bands_contig = 1;
    for (int s = 1; s < cfg->nsegs; s++)
      if (cfg->segs[s].band < cfg->segs[s - 1].band) b.bands_contig = 0;
    int s = 0;
    for (int band = 0; band <= cfg->nbands && band <= CRN_MAX_BANDS; band++) {
      while (b.bands_contig && s < cfg->nsegs && cfg->segs[s].band < band) s++;
      b.band_first[band] = (short)s;
    }
  }
  {  // spectrum slices (accumulator registers) the band table reads: selects the pruned kernel when it can
    const crn::RadixPlan rp = crn::radix_plan(cfg->nfft);
    const int per_slice = cfg->nfft / rp.e;
    const char *full = getenv("CRN_NO_PRUNE");  // development override: always the all-bins kernel
    b.acc_mask = (full && full[0] == '1') ? 0xFFFFFFFFu : 0u;
    for (int s = 0; s < cfg->nsegs; s++)
      for (int k = cfg->segs[s].lo; k < cfg->segs[s].hi; k++) b.acc_mask |= 1u << (k / per_slice);
  }

  st = h->launch(b, cfg->window, cfg->detector, 0, nullptr, &h->geo);
  if (st != CRN_OK) return st;
  if (h->geo.ctas_per_sm < 1) {
    return crn::fail(CRN_ERR_CUDA, "kernel %s does not fit on an SM (smem %d B)", h->geo.name, h->geo.smem_bytes);
  }
  // Epilogue strategy (crn_sense_kernel.cuh): CTA-wide when every team gets the same, long run of frames
  // per group; otherwise the barrier-free unit epilogue with the best-balanced units-per-group.
  {
    const int teams = h->geo.teams;
    const bool long_even_groups = (cfg->navg % teams == 0) && (cfg->navg / teams >= 8);
    const char *no_tma = getenv("CRN_NO_TMA");
    h->allow_tma = !(no_tma && no_tma[0] == '1');
    const char *rc = getenv("CRN_RING_COPY");
    h->ring_zero_copy = (rc && (rc[0] == '0' || rc[0] == '1'))
                            ? rc[0] == '0'
                            : h->sample_bytes * (size_t)cfg->navg * cfg->frame_len >= ((size_t)2 << 20);
    const char *cap = getenv("CRN_SPLIT");  // development override: largest split (1 = never split a group)
    if (cap && atoi(cap) >= 1) h->max_split = atoi(cap);
    const char *fsp = getenv("CRN_FORCE_SPLIT");
    if (fsp && atoi(fsp) >= 1) h->force_split = atoi(fsp);
    const char *epi = getenv("CRN_EPILOGUE_FRAMES");
    if (epi && atof(epi) >= 0.0) h->epilogue_frames = atof(epi);
    const char *force = getenv("CRN_EPI");  // "cta" | "unit": development override
    bool cta = long_even_groups;
    if (force && !strcmp(force, "cta")) cta = true;
    if (force && !strcmp(force, "unit")) cta = false;
    b.upg = cta ? 0 : pick_upg(h->geo.units, h->geo.teams_per_unit, cfg->navg);
    st = h->launch(b, cfg->window, cfg->detector, 0, nullptr, &h->geo);  // attributes of the chosen variant
    if (st != CRN_OK) return st;
  }

  CRN_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CRN_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  // split scratch of the two batch paths, sized once: splitting only pays while a batch leaves CTAs idle, so batches
  // beyond two full grids are never split
  h->split_max_groups = 2 * (int64_t)h->num_sms * (h->geo.ctas_per_sm > 0 ? h->geo.ctas_per_sm : 1);
  if (b.upg == 0) {
    st = ensure_split_buffers(h, h->dev_split, h->split_max_groups, h->max_split);
    if (st != CRN_OK) return st;
    st = ensure_split_buffers(h, h->host_split, h->split_max_groups, h->max_split);
    if (st != CRN_OK) return st;
  }
  CRN_CUDA(cudaEventCreateWithFlags(&h->dev_split_event, cudaEventDisableTiming));

  // streaming ring: ring_slots decisions of K frames each
  const size_t slot_bytes = h->sample_bytes * (size_t)cfg->navg * cfg->frame_len;
  h->ring.resize(h->cfg.ring_slots);
  for (auto &s : h->ring) {
    CRN_CUDA(cudaMallocHost(&s.h_iq, slot_bytes));
    CRN_CUDA(cudaMalloc(&s.d_iq, slot_bytes));
    st = alloc_results(s.res, 1, cfg->nbands);
    if (st != CRN_OK) return st;
    st = ensure_split_buffers(h, s.split, 1, h->max_split);  // now, not in front of the first decision
    if (st != CRN_OK) return st;
    CRN_CUDA(cudaHostGetDevicePointer((void **)&s.z_iq, s.h_iq, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&s.z_feat, s.res.h_feat, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&s.z_ann, s.res.h_ann, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&s.z_dec, s.res.h_dec, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&s.z_mask, s.res.h_mask, 0));
    CRN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    s.rv = &s.res;
    s.ri = 0;
  }
  guard.h = nullptr;
  *out = h;
  return CRN_OK;
}

namespace {
void destroy_pool(Pool *p) {
  if (!p) return;
  if (p->proto) {
    cudaSetDevice(p->proto->device);
    if (p->proto->stream) cudaStreamSynchronize(p->proto->stream);
  }
  for (auto &ps : p->slots) {
    cudaFreeHost(ps.h_iq);
    cudaFree(ps.d_iq);
    free_results(ps.res);
    free_split_buffers(ps.split);
    if (ps.done) cudaEventDestroy(ps.done);
  }
  if (p->proto) crn_destroy(p->proto);
  delete p;
}
}  // namespace

int crn_create_many(const crn_config *cfg, int32_t n, crn_handle **out) {
  if (!out || n < 1) return crn::fail(CRN_ERR_INVALID, "crn_create_many: bad argument");
  for (int i = 0; i < n; i++) out[i] = nullptr;
  Pool *p = new (std::nothrow) Pool();
  if (!p) return crn::fail(CRN_ERR_NOMEM, "out of host memory");
  struct Guard {
    Pool *p;
    ~Guard() {
      if (!p) return;
      for (crn_handle *m : p->members) delete m;
      destroy_pool(p);
    }
  } guard{p};
  int st = crn_create(cfg, &p->proto);
  if (st != CRN_OK) return st;
  crn_handle *pr = p->proto;
  p->n = n;
  p->slot_bytes = pr->sample_bytes * (size_t)pr->cfg.navg * pr->cfg.frame_len;
  const int nbands = pr->cfg.nbands;
  const size_t per_member_scratch = (size_t)pr->max_split * pr->cfg.nsegs;
  p->slots.resize(pr->ring.size());
  for (auto &ps : p->slots) {
    CRN_CUDA(cudaMallocHost(&ps.h_iq, p->slot_bytes * n));
    CRN_CUDA(cudaMalloc(&ps.d_iq, p->slot_bytes * n));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&ps.z_iq, ps.h_iq, 0));
    ps.res.cap = n;
    CRN_CUDA(cudaMallocHost(&ps.res.h_feat, sizeof(float) * n * nbands));
    CRN_CUDA(cudaMallocHost(&ps.res.h_ann, sizeof(double) * n * 3));
    CRN_CUDA(cudaMallocHost(&ps.res.h_dec, sizeof(int32_t) * n));
    CRN_CUDA(cudaMallocHost(&ps.res.h_mask, sizeof(unsigned long long) * n));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&ps.z_feat, ps.res.h_feat, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&ps.z_ann, ps.res.h_ann, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&ps.z_dec, ps.res.h_dec, 0));
    CRN_CUDA(cudaHostGetDevicePointer((void **)&ps.z_mask, ps.res.h_mask, 0));
    CRN_CUDA(cudaMalloc(&ps.split.d_scratch, sizeof(float) * per_member_scratch * n));
    ps.split.scratch_cap = per_member_scratch * n;
    CRN_CUDA(cudaMalloc(&ps.split.d_gcount, sizeof(int) * n));
    CRN_CUDA(cudaMemset(ps.split.d_gcount, 0, sizeof(int) * n));
    ps.split.gcount_cap = (size_t)n;
    CRN_CUDA(cudaEventCreateWithFlags(&ps.done, cudaEventDisableTiming));
  }
  for (int i = 0; i < n; i++) {
    crn_handle *m = new (std::nothrow) crn_handle();
    if (!m) return crn::fail(CRN_ERR_NOMEM, "out of host memory");
    p->members.push_back(m);
    m->cfg = pr->cfg;
    m->pool = p;
    m->pool_index = i;
    m->device = pr->device;
    m->num_sms = pr->num_sms;
    m->stride = pr->stride;
    m->sample_bytes = pr->sample_bytes;
    m->allow_tma = pr->allow_tma;
    m->launch = pr->launch;
    m->geo = pr->geo;
    m->base = pr->base;          // table pointers are the prototype's
    m->max_split = pr->max_split;
    m->force_split = pr->force_split;
    m->epilogue_frames = pr->epilogue_frames;
    m->ring_zero_copy = pr->ring_zero_copy;
    m->split_max_groups = pr->split_max_groups;  // (members have no batch scratch: batch calls on a member are refused)
    m->stream = pr->stream;       // one stream for the pool: member launches and pooled launches are ordered
    m->ring.resize(p->slots.size());
    for (size_t k = 0; k < m->ring.size(); k++) {
      RingSlot &s = m->ring[k];
      Pool::PSlot &ps = p->slots[k];
      s.h_iq = ps.h_iq + p->slot_bytes * i;
      s.d_iq = ps.d_iq + p->slot_bytes * i;
      s.z_iq = reinterpret_cast<const float2 *>(reinterpret_cast<const unsigned char *>(ps.z_iq) + p->slot_bytes * i);
      s.z_feat = ps.z_feat + (size_t)i * nbands;
      s.z_ann = ps.z_ann + (size_t)i * 3;
      s.z_dec = ps.z_dec + i;
      s.z_mask = ps.z_mask + i;
      s.split.d_scratch = ps.split.d_scratch + per_member_scratch * i;
      s.split.scratch_cap = per_member_scratch;
      s.split.d_gcount = ps.split.d_gcount + i;
      s.split.gcount_cap = 1;
      s.rv = &ps.res;
      s.ri = i;
      CRN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
  }
  p->live = n;
  for (int i = 0; i < n; i++) out[i] = p->members[i];
  guard.p = nullptr;
  return CRN_OK;
}

int crn_destroy(crn_handle *h) {
  if (!h) return CRN_OK;
  if (h->pool) {  // a member owns nothing but its events; the last member to go takes the pool with it
    Pool *p = h->pool;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &s : h->ring)
      if (s.done) cudaEventDestroy(s.done);
    delete h;
    if (--p->live == 0) destroy_pool(p);
    return CRN_OK;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (auto &s : h->ring) {
    cudaFreeHost(s.h_iq);
    cudaFree(s.d_iq);
    free_results(s.res);
    free_split_buffers(s.split);
    if (s.done) cudaEventDestroy(s.done);
  }
  free_staging(h);
  if (h->dev_split_event) cudaEventDestroy(h->dev_split_event);
  cudaFree(h->d_tw);
  cudaFree(h->d_win);
  free_split_buffers(h->host_split);
  free_split_buffers(h->dev_split);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return CRN_OK;
}

// ---- streaming path -------------------------------------------------------------------------------

int crn_ring_acquire(crn_handle *h, void **slot) {
  if (!h || !slot) return crn::fail(CRN_ERR_INVALID, "crn_ring_acquire: null argument");
  RingSlot &s = h->ring[h->fill_slot];
  if (s.state != 0) return crn::fail(CRN_ERR_OVERRUN, "ring full: %d decisions unread", h->inflight);
  if (h->fill_frames >= h->cfg.navg)  // cannot happen through this API (crn_submit rolls back on failure); never hand out
    return crn::fail(CRN_ERR_INVALID, "crn_ring_acquire: slot already holds %d frames", h->fill_frames);  // memory past the slot
  *slot = s.h_iq + h->sample_bytes * (size_t)h->fill_frames * h->cfg.frame_len;
  return CRN_OK;
}

int crn_submit(crn_handle *h, int32_t nframes) {
  if (!h || nframes < 1) return crn::fail(CRN_ERR_INVALID, "crn_submit: bad argument");
  if (h->fill_frames + nframes > h->cfg.navg)
    return crn::fail(CRN_ERR_INVALID, "crn_submit: %d frames would cross a decision boundary", nframes);
  RingSlot &s = h->ring[h->fill_slot];
  if (s.state != 0) return crn::fail(CRN_ERR_OVERRUN, "ring full: %d decisions unread", h->inflight);
  if (h->fill_frames == 0) s.first_frame = h->frames_seen;
  if (h->fill_frames + nframes < h->cfg.navg) {
    h->fill_frames += nframes;
    h->frames_seen += (uint64_t)nframes;
    return CRN_OK;
  }
  // K-th frame: ship the slot and enqueue the fused kernel (stream ordered, non blocking).  The frame counters are
  // committed only once the launch is queued: a CUDA failure leaves the slot as it was before this call, so the next
  // crn_ring_acquire stays inside the slot (the caller resubmits or calls crn_reset).
  CRN_CUDA(cudaSetDevice(h->device));
  const size_t slot_bytes = h->sample_bytes * (size_t)h->cfg.navg * h->cfg.frame_len;
  crn::SenseParams p = h->base;
  p.stride = h->cfg.frame_len;  // ring slots are packed
  if (h->ring_zero_copy) {
    // every sample is read exactly once, so the kernel can pull it across PCIe itself: transfer and FFTs overlap
    // frame by frame and there is no copy -> launch hand-over on the stream
    p.iq = s.z_iq;
  } else {
    CRN_CUDA(cudaMemcpyAsync(s.d_iq, s.h_iq, slot_bytes, cudaMemcpyHostToDevice, h->stream));
    p.iq = reinterpret_cast<const float2 *>(s.d_iq);
  }
  // the decision's 52..300 bytes are written by the kernel straight into the slot's page-locked host mirror
  // (pinned memory is device-addressable): no device->host copies queue up behind the kernel
  p.feat = s.z_feat;
  p.ann = s.z_ann;
  p.decision = s.z_dec;
  p.mask = s.z_mask;
  p.ngroups = 1;
  int grid = 1;
  int st = shape_launch(h, p, 1, s.split, &grid);  // one decision: its K frames are dealt to several CTAs
  if (st != CRN_OK) return st;
  st = h->launch(p, h->cfg.window, h->cfg.detector, grid, h->stream, nullptr);
  if (st != CRN_OK) return st;
  h->launches++;
  CRN_CUDA(cudaEventRecord(s.done, h->stream));
  h->frames_seen += (uint64_t)nframes;
  s.state = 1;
  h->inflight++;
  h->fill_slot = (h->fill_slot + 1) % (int)h->ring.size();
  h->fill_frames = 0;
  return CRN_OK;
}

int crn_submit_many(crn_handle *const *hs, int32_t n, int32_t nframes) {
  if (!hs || n < 1 || nframes < 1) return crn::fail(CRN_ERR_INVALID, "crn_submit_many: bad argument");
  for (int i = 0; i < n; i++)
    if (!hs[i]) return crn::fail(CRN_ERR_INVALID, "crn_submit_many: null handle");
  // One launch needs the members of one pool, all of them, all at the same point of the same slot; anything else is
  // served handle by handle (same results, n launches).
  Pool *p = hs[0]->pool;
  bool poolable = p && n == p->n;
  if (poolable) {
    std::vector<char> seen((size_t)n, 0);
    for (int i = 0; i < n && poolable; i++) {
      crn_handle *h = hs[i];
      poolable = h->pool == p && !seen[h->pool_index] && h->fill_slot == hs[0]->fill_slot &&
                 h->fill_frames == hs[0]->fill_frames;
      if (poolable) seen[h->pool_index] = 1;
    }
  }
  if (!poolable) {
    for (int i = 0; i < n; i++) {
      int st = crn_submit(hs[i], nframes);
      if (st != CRN_OK) return st;
    }
    return CRN_OK;
  }
  crn_handle *pr = p->proto;
  const int K = pr->cfg.navg, slot = hs[0]->fill_slot, fill = hs[0]->fill_frames;
  if (fill + nframes > K)
    return crn::fail(CRN_ERR_INVALID, "crn_submit_many: %d frames would cross a decision boundary", nframes);
  for (int i = 0; i < n; i++)
    if (hs[i]->ring[slot].state != 0) return crn::fail(CRN_ERR_OVERRUN, "ring full: %d decisions unread", hs[i]->inflight);
  if (fill + nframes < K) {
    for (int i = 0; i < n; i++) {
      if (fill == 0) hs[i]->ring[slot].first_frame = hs[i]->frames_seen;
      hs[i]->fill_frames += nframes;
      hs[i]->frames_seen += (uint64_t)nframes;
    }
    return CRN_OK;
  }
  // K-th frame of every member: one copy, one launch over n decision groups ([radio][K][L] is a batch of n groups)
  CRN_CUDA(cudaSetDevice(pr->device));
  Pool::PSlot &ps = p->slots[slot];
  crn::SenseParams prm = pr->base;
  prm.stride = pr->cfg.frame_len;  // ring slots are packed
  if (pr->ring_zero_copy) {
    prm.iq = ps.z_iq;
  } else {
    CRN_CUDA(cudaMemcpyAsync(ps.d_iq, ps.h_iq, p->slot_bytes * n, cudaMemcpyHostToDevice, pr->stream));
    prm.iq = reinterpret_cast<const float2 *>(ps.d_iq);
  }
  prm.feat = ps.z_feat;
  prm.ann = ps.z_ann;
  prm.decision = ps.z_dec;
  prm.mask = ps.z_mask;
  prm.ngroups = n;
  prm.use_tma = !pr->ring_zero_copy && (prm.upg == 0) && pr->allow_tma && ((reinterpret_cast<uintptr_t>(prm.iq) & 15) == 0) &&
                ((pr->cfg.frame_len * pr->sample_bytes) % 16 == 0);
  int grid = 1;
  int st = shape_launch(pr, prm, n, ps.split, &grid);  // few radios: their K frames are dealt to several CTAs each
  if (st != CRN_OK) return st;
  st = pr->launch(prm, pr->cfg.window, pr->cfg.detector, grid, pr->stream, nullptr);
  if (st != CRN_OK) return st;
  pr->launches++;
  CRN_CUDA(cudaEventRecord(ps.done, pr->stream));
  for (int i = 0; i < n; i++) {
    crn_handle *h = hs[i];
    RingSlot &s = h->ring[slot];
    if (fill == 0) s.first_frame = h->frames_seen;
    h->frames_seen += (uint64_t)nframes;
    s.pool_done = ps.done;
    s.state = 1;
    h->inflight++;
    h->fill_slot = (h->fill_slot + 1) % (int)h->ring.size();
    h->fill_frames = 0;
  }
  return CRN_OK;
}

static int take_result(crn_handle *h, crn_result *out, bool block) {
  if (!h || !out) return crn::fail(CRN_ERR_INVALID, "crn_poll/wait: null argument");
  if (h->inflight == 0) return crn::fail(CRN_ERR_NOT_READY, "no decision in flight");
  RingSlot &s = h->ring[h->tail_slot];
  cudaEvent_t done = s.pool_done ? s.pool_done : s.done;
  if (block) {
    CRN_CUDA(cudaEventSynchronize(done));
  } else {
    cudaError_t e = cudaEventQuery(done);
    if (e == cudaErrorNotReady) return CRN_ERR_NOT_READY;
    if (e != cudaSuccess) return crn::fail(CRN_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
  }
  unpack_result(*s.rv, s.ri, h->cfg.nbands, s.first_frame, out);
  s.pool_done = nullptr;
  s.state = 0;
  h->inflight--;
  h->tail_slot = (h->tail_slot + 1) % (int)h->ring.size();
  return CRN_OK;
}
int crn_poll(crn_handle *h, crn_result *out) { return take_result(h, out, false); }
int crn_wait(crn_handle *h, crn_result *out) { return take_result(h, out, true); }

int crn_reset(crn_handle *h) {
  if (!h) return crn::fail(CRN_ERR_INVALID, "crn_reset: null handle");
  h->fill_frames = 0;
  return CRN_OK;
}

// ---- batch paths ----------------------------------------------------------------------------------

int crn_sense_batch_device(crn_handle *h, const void *d_iq, int64_t ngroups, float *d_feat,
                           double *d_ann, int32_t *d_decision, uint64_t *d_mask, void *cuda_stream) {
  if (!h || !d_iq || !d_feat || ngroups < 0)
    return crn::fail(CRN_ERR_INVALID, "crn_sense_batch_device: bad argument");
  if (h->pool) return crn::fail(CRN_ERR_UNSUPPORTED, "batch calls need a handle from crn_create, not a crn_create_many member");
  CRN_CUDA(cudaSetDevice(h->device));
  return launch(h, h->dev_split, d_iq, ngroups, d_feat, d_ann, d_decision,
                (unsigned long long *)d_mask, (cudaStream_t)cuda_stream);
}

int crn_sense_batch_host(crn_handle *h, const void *iq_, int64_t ngroups, crn_result *results) {
  const unsigned char *iq = static_cast<const unsigned char *>(iq_);
  if (!h || !iq || !results || ngroups < 0)
    return crn::fail(CRN_ERR_INVALID, "crn_sense_batch_host: bad argument");
  if (h->pool) return crn::fail(CRN_ERR_UNSUPPORTED, "batch calls need a handle from crn_create, not a crn_create_many member");
  if (ngroups == 0) return CRN_OK;
  CRN_CUDA(cudaSetDevice(h->device));
  const size_t group_bytes = h->sample_bytes * (size_t)h->stride * h->cfg.navg;
  if (h->chunk_groups == 0) {
    // ~64 MiB chunks: large enough to amortise launch + copy latency, small enough to overlap
    int64_t cg = (int64_t)((64u << 20) / group_bytes);
    if (cg < 1) cg = 1;
    // chunk_groups is set only when every staging buffer exists; a failure part-way releases what was allocated, so
    // the next call starts over instead of running on null buffers
    auto stage_alloc = [&]() -> int {
      for (int i = 0; i < 2; i++) {
        CRN_CUDA(cudaMallocHost(&h->h_stage[i], cg * group_bytes));
        CRN_CUDA(cudaMalloc(&h->d_stage[i], cg * group_bytes));
        int st = alloc_results(h->stage_res[i], cg, h->cfg.nbands);
        if (st != CRN_OK) return st;
        CRN_CUDA(cudaEventCreateWithFlags(&h->stage_done[i], cudaEventDisableTiming));
        CRN_CUDA(cudaEventCreateWithFlags(&h->stage_copied[i], cudaEventDisableTiming));
      }
      return CRN_OK;
    };
    const int ast = stage_alloc();
    if (ast != CRN_OK) {
      free_staging(h);
      return ast;
    }
    h->chunk_groups = cg;
  }
  // Is the caller's buffer already page-locked?  Then DMA straight from it.
  cudaPointerAttributes pa;
  bool pinned = false;
  if (cudaPointerGetAttributes(&pa, iq) == cudaSuccess) pinned = (pa.type == cudaMemoryTypeHost);
  else cudaGetLastError();

  const int64_t cg = h->chunk_groups;
  const int64_t nchunks = (ngroups + cg - 1) / cg;
  int64_t pending_base[2] = {-1, -1};
  int64_t pending_n[2] = {0, 0};
  auto drain = [&](int b) -> int {
    if (pending_base[b] < 0) return CRN_OK;
    CRN_CUDA(cudaEventSynchronize(h->stage_done[b]));
    for (int64_t i = 0; i < pending_n[b]; i++)
      unpack_result(h->stage_res[b], i, h->cfg.nbands,
                    (uint64_t)(pending_base[b] + i) * (uint64_t)h->cfg.navg, &results[pending_base[b] + i]);
    pending_base[b] = -1;
    return CRN_OK;
  };
  for (int64_t c = 0; c < nchunks; c++) {
    const int b = (int)(c & 1);
    int st = drain(b);  // buffer b free again (its kernel and readback finished)
    if (st != CRN_OK) return st;
    const int64_t g0 = c * cg;
    const int64_t n = (ngroups - g0 < cg) ? (ngroups - g0) : cg;
    const unsigned char *src = iq + g0 * group_bytes;
    if (!pinned) {
      memcpy(h->h_stage[b], src, n * group_bytes);
      src = h->h_stage[b];
    }
    // H2D on its own stream so the next chunk's copy overlaps this chunk's kernel and result readback
    CRN_CUDA(cudaMemcpyAsync(h->d_stage[b], src, n * group_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CRN_CUDA(cudaEventRecord(h->stage_copied[b], h->copy_stream));
    CRN_CUDA(cudaStreamWaitEvent(h->stream, h->stage_copied[b], 0));
    ResultBuf &r = h->stage_res[b];
    st = launch(h, h->host_split, h->d_stage[b], n, r.d_feat, r.d_ann, r.d_dec, r.d_mask, h->stream);
    if (st != CRN_OK) return st;
    st = fetch_results_async(r, n, h->cfg.nbands, h->stream);
    if (st != CRN_OK) return st;
    CRN_CUDA(cudaEventRecord(h->stage_done[b], h->stream));
    pending_base[b] = g0;
    pending_n[b] = n;
  }
  int st = drain((int)(nchunks & 1));
  if (st != CRN_OK) return st;
  return drain((int)((nchunks + 1) & 1));
}

// a pool member reports its own launches plus the pool's shared ones (crn_submit_many)
int64_t crn_launch_count(const crn_handle *h) {
  if (!h) return 0;
  return h->launches + (h->pool ? h->pool->proto->launches : 0);
}

int crn_get_kernel_info(const crn_handle *h, crn_kernel_info *info) {
  if (!h || !info) return crn::fail(CRN_ERR_INVALID, "crn_get_kernel_info: null argument");
  memset(info, 0, sizeof(*info));
  info->nfft = h->cfg.nfft;
  info->threads_per_frame = h->geo.threads_per_frame;
  info->elems_per_thread = h->geo.elems_per_thread;
  info->teams_per_cta = h->geo.teams;
  info->threads_per_cta = h->geo.threads_per_cta;
  info->ctas_per_sm = h->geo.ctas_per_sm;
  info->grid = h->num_sms * h->geo.ctas_per_sm;
  info->smem_bytes = h->geo.smem_bytes;
  info->regs_per_thread = h->geo.regs;
  info->num_sms = h->num_sms;
  snprintf(info->name, sizeof(info->name), "%s", h->geo.name);
  return CRN_OK;
}

}  // extern "C"
