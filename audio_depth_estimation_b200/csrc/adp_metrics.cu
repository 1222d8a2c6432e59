// Per-sample depth error metrics of a whole batch in three launches, results stay on the device.
//
// Replaces the per-sample D2H copy + numpy loop of the reference's validation / test loops:
//   train.py:795-838, test.py:231-285   (x max_depth when depth_norm, clip pred to [eps, max_depth], gt >= 0)
//   utils_criterion.py:6-90             (compute_errors: abs_rel, rmse, a1, a2, a3, log_10, mae)
//
// compute_errors picks its epsilon from max(gt) and falls back to a second pixel set when no prediction exceeds
// it (:38-54), so the batch is read twice: pass 1 finds max(gt | gt != 0) per sample, pass 2 accumulates the
// sums of the primary set V = {gt != 0, pred > e, gt > e} and of the fall-back set W = {gt != 0, gt > e, pred > 0}
// (for both values the second epsilon can take in W), pass 3 selects the branch and forms the seven numbers.
// Element-wise arithmetic is fp32 like the reference's numpy code (the threshold tests are bit-exact), the
// accumulation is fp64.
#include <math.h>
#include <algorithm>
#include "adp_common.cuh"

namespace {

constexpr int MET_THREADS = 256;
constexpr int MET_BLOCKS_PER_SAMPLE = 32;

// accumulator slots per sample (doubles)
enum {
  V_CNT = 0, V_ABSREL, V_SQ, V_MAE, V_A1, V_A2, V_A3, V_LOG,
  W_CNT, W_ABSREL, W_SQ, W_MAE,
  W_A1_HI, W_A2_HI, W_A3_HI, W_LOG_HI,      // eps2 = 1e-3
  W_A1_LO, W_A2_LO, W_A3_LO, W_LOG_LO,      // eps2 = 1e-6
  G_CNT,                                    // gt > e (any prediction)
  MET_SLOTS
};
// int slots per sample: ordered max(gt | gt != 0), count(gt != 0), ordered max(gt | W)
constexpr int MET_INTS = 4;

struct Prep {
  float scale, clip_lo, clip_hi;
  int prepare;
};

__device__ __forceinline__ void prep(const Prep& c, float& g, float& p) {
  g *= c.scale;
  p *= c.scale;
  if (c.prepare) {
    p = fminf(fmaxf(p, c.clip_lo), c.clip_hi);
    g = fmaxf(g, 0.f);
  }
}

__global__ void metrics_init_kernel(double* acc, int* ints, int batch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch * MET_SLOTS) acc[i] = 0.0;
  if (i < batch) {
    ints[i * MET_INTS + 0] = float_to_ordered(-INFINITY);
    ints[i * MET_INTS + 1] = 0;
    ints[i * MET_INTS + 2] = float_to_ordered(-INFINITY);
    ints[i * MET_INTS + 3] = 0;
  }
}

__global__ void __launch_bounds__(MET_THREADS)
metrics_gmax_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, Prep c, int* __restrict__ ints) {
  const int b = blockIdx.y;
  const float* g0 = gt + (size_t)b * n;
  const float* p0 = pred + (size_t)b * n;
  float gmax = -INFINITY;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * MET_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MET_THREADS) {
    float g = g0[i], p = p0[i];
    prep(c, g, p);
    if (g != 0.f) {
      gmax = fmaxf(gmax, g);
      ++cnt;
    }
  }
  gmax = warp_max(gmax);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicMax(&ints[b * MET_INTS + 0], float_to_ordered(gmax));
    atomicAdd(&ints[b * MET_INTS + 1], cnt);
  }
}

__global__ void __launch_bounds__(MET_THREADS)
metrics_accum_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, Prep c,
                     double* __restrict__ acc, int* __restrict__ ints) {
  const int b = blockIdx.y;
  if (ints[b * MET_INTS + 1] == 0) return;                       // no valid ground truth: zeros (:25-27)
  const float e = ordered_to_float(ints[b * MET_INTS + 0]) > 1.0f ? 1e-3f : 1e-6f;
  const float* g0 = gt + (size_t)b * n;
  const float* p0 = pred + (size_t)b * n;
  float s[MET_SLOTS];
#pragma unroll
  for (int k = 0; k < MET_SLOTS; ++k) s[k] = 0.f;
  float wmax = -INFINITY;
  // fp32 partial sums per thread are short (n / (32 * 256) elements); they are widened before the reduction
  for (long long i = (long long)blockIdx.x * MET_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MET_THREADS) {
    float g = g0[i], p = p0[i];
    prep(c, g, p);
    if (g == 0.f || !(g > e)) continue;
    s[G_CNT] += 1.f;
    const float diff = fabsf(g - p);
    if (p > e) {                                                   // primary set: every clamp of :60-83 is a no-op
      const float th = fmaxf(g / p, p / g);
      s[V_CNT] += 1.f;
      s[V_ABSREL] += diff / g;
      s[V_SQ] += (g - p) * (g - p);
      s[V_MAE] += diff;
      s[V_A1] += th < 1.25f ? 1.f : 0.f;
      s[V_A2] += th < 1.5625f ? 1.f : 0.f;
      s[V_A3] += th < 1.953125f ? 1.f : 0.f;
      s[V_LOG] += fabsf(log10f(g) - log10f(p));
    } else if (p > 0.f) {                                          // fall-back set (:42-47); pred <= e here
      wmax = fmaxf(wmax, g);
      s[W_CNT] += 1.f;
      s[W_ABSREL] += diff / g;
      s[W_SQ] += (g - p) * (g - p);
      s[W_MAE] += diff;
      {
        const float pc = fmaxf(p, 1e-3f), th = fmaxf(g / pc, pc / g);
        s[W_A1_HI] += th < 1.25f ? 1.f : 0.f;
        s[W_A2_HI] += th < 1.5625f ? 1.f : 0.f;
        s[W_A3_HI] += th < 1.953125f ? 1.f : 0.f;
        s[W_LOG_HI] += fabsf(log10f(fmaxf(g, 1e-3f)) - log10f(pc));
      }
      {
        const float pc = fmaxf(p, 1e-6f), th = fmaxf(g / pc, pc / g);
        s[W_A1_LO] += th < 1.25f ? 1.f : 0.f;
        s[W_A2_LO] += th < 1.5625f ? 1.f : 0.f;
        s[W_A3_LO] += th < 1.953125f ? 1.f : 0.f;
        s[W_LOG_LO] += fabsf(log10f(fmaxf(g, 1e-6f)) - log10f(pc));
      }
    }
  }
  __shared__ double red[MET_THREADS / 32][MET_SLOTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < MET_SLOTS; ++k) {
    double d = warp_sum((double)s[k]);
    if (lane == 0) red[warp][k] = d;
  }
  wmax = warp_max(wmax);
  if (lane == 0 && wmax > -INFINITY) atomicMax(&ints[b * MET_INTS + 2], float_to_ordered(wmax));
  __syncthreads();
  if (threadIdx.x < MET_SLOTS) {
    double t = 0.0;
    for (int w = 0; w < MET_THREADS / 32; ++w) t += red[w][threadIdx.x];
    if (t != 0.0) atomicAdd(&acc[(size_t)b * MET_SLOTS + threadIdx.x], t);
  }
}

__device__ __forceinline__ double clean(double v) { return (v != v || isinf(v)) ? 0.0 : v; }

// out[b] = (abs_rel, rmse, a1, a2, a3, log_10, mae) -- the reference's return order (:90)
__global__ void metrics_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ ints, int batch,
                                        double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* a = acc + (size_t)b * MET_SLOTS;
  double* o = out + (size_t)b * 7;
  for (int k = 0; k < 7; ++k) o[k] = 0.0;
  if (ints[b * MET_INTS + 1] == 0) return;
  const double gmax = (double)ordered_to_float(ints[b * MET_INTS + 0]);
  if (a[V_CNT] > 0.0) {
    const double n = a[V_CNT];
    o[0] = clean(a[V_ABSREL] / n); o[1] = clean(sqrt(a[V_SQ] / n));
    o[2] = clean(a[V_A1] / n); o[3] = clean(a[V_A2] / n); o[4] = clean(a[V_A3] / n);
    o[5] = clean(a[V_LOG] / n); o[6] = clean(a[V_MAE] / n);
    return;
  }
  if (a[G_CNT] == 0.0) return;                                    // no gt above epsilon (:43-45)
  if (a[W_CNT] == 0.0) {                                          // every prediction <= 0 (:48-54)
    o[0] = 1.0; o[1] = gmax; o[5] = 1.0; o[6] = gmax;
    return;
  }
  const double n = a[W_CNT];
  const bool hi = ordered_to_float(ints[b * MET_INTS + 2]) > 1.0f;
  o[0] = clean(a[W_ABSREL] / n); o[1] = clean(sqrt(a[W_SQ] / n));
  o[2] = clean(a[hi ? W_A1_HI : W_A1_LO] / n); o[3] = clean(a[hi ? W_A2_HI : W_A2_LO] / n);
  o[4] = clean(a[hi ? W_A3_HI : W_A3_LO] / n);
  o[5] = clean(a[hi ? W_LOG_HI : W_LOG_LO] / n); o[6] = clean(a[W_MAE] / n);
}

}  // namespace

extern "C" size_t adp_depth_metrics_workspace_bytes(int batch) {
  if (batch <= 0) return 0;
  return adp_align_up((size_t)batch * MET_SLOTS * sizeof(double), 256) + adp_align_up((size_t)batch * MET_INTS * sizeof(int), 256);
}

extern "C" int adp_depth_metrics(const float* pred, const float* gt, int batch, int64_t n_per_sample, float scale,
                                 int prepare, float clip_lo, float clip_hi, double* metrics, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  ADP_CHECK_ARG(pred && gt && metrics && workspace, "depth_metrics: null pointer");
  ADP_CHECK_ARG(batch > 0 && batch <= 65535 && n_per_sample > 0, "depth_metrics: bad batch/n (%d, %lld)", batch,
                (long long)n_per_sample);
  ADP_CHECK_ARG(workspace_bytes >= adp_depth_metrics_workspace_bytes(batch), "depth_metrics: workspace too small (%zu)",
                workspace_bytes);
  ADP_CHECK_ARG(!prepare || clip_hi >= clip_lo, "depth_metrics: clip range");
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = reinterpret_cast<double*>(workspace);
  int* ints = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) +
                                     adp_align_up((size_t)batch * MET_SLOTS * sizeof(double), 256));
  Prep c{scale, clip_lo, clip_hi, prepare ? 1 : 0};
  metrics_init_kernel<<<adp_cdiv(batch * MET_SLOTS, 256), 256, 0, s>>>(acc, ints, batch);
  ADP_LAUNCH_CHECK();
  const int per = (int)std::min<long long>(MET_BLOCKS_PER_SAMPLE, adp_cdiv((long long)n_per_sample, (long long)MET_THREADS));
  dim3 grid(per, batch);
  metrics_gmax_kernel<<<grid, MET_THREADS, 0, s>>>(pred, gt, (long long)n_per_sample, c, ints);
  ADP_LAUNCH_CHECK();
  metrics_accum_kernel<<<grid, MET_THREADS, 0, s>>>(pred, gt, (long long)n_per_sample, c, acc, ints);
  ADP_LAUNCH_CHECK();
  metrics_finalize_kernel<<<adp_cdiv(batch, 128), 128, 0, s>>>(acc, ints, batch, metrics);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}
