// Masked depth loss (L1 + SIlog), forward statistics and fused gradient.
//
// Replaces the boolean-index gathers and elementwise/reduction chain of the reference at
//   train.py:646-669   (mask gt != 0, optional x max_depth, Combined = l1_w*L1 + silog_w*SIlog)
//   utils_loss.py:29-49 (clamp(min=eps), d = log p - log g, sqrt(clamp(mean(d^2) - lam*mean(d)^2, 0)))
// The loss over the valid pixels of the (global) batch is a function of four sums
//   {N, sum|p-g|, sum d, sum d^2}; they are produced by one pass (adp_depth_loss_sums), can be
// all-reduced across data-parallel ranks, and the gradient pass (adp_depth_loss_backward)
// reads them back: no gather, no host synchronisation.
#include "adp_common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ void loss_accum(float p, float g, float scale, float eps, int use_mask, float& n,
                                           float& sa, float& sd, float& sd2) {
  if (!use_mask || g != 0.0f) {
    float ps = p * scale, gs = g * scale;
    float d = logf(fmaxf(ps, eps)) - logf(fmaxf(gs, eps));
    n += 1.f;
    sa += fabsf(ps - gs);
    sd += d;
    sd2 += d * d;
  }
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, float scale,
                 float eps, int use_mask, double* __restrict__ sums) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = ld4(pred + 4 * i), g = ld4(gt + 4 * i);
    float c = 0.f, sa = 0.f, sd = 0.f, sd2 = 0.f;
    loss_accum(p.x, g.x, scale, eps, use_mask, c, sa, sd, sd2);
    loss_accum(p.y, g.y, scale, eps, use_mask, c, sa, sd, sd2);
    loss_accum(p.z, g.z, scale, eps, use_mask, c, sa, sd, sd2);
    loss_accum(p.w, g.w, scale, eps, use_mask, c, sa, sd, sd2);
    acc[0] += c; acc[1] += sa; acc[2] += sd; acc[3] += sd2;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 << 2; i < n; ++i) {
      float c = 0.f, sa = 0.f, sd = 0.f, sd2 = 0.f;
      loss_accum(pred[i], gt[i], scale, eps, use_mask, c, sa, sd, sd2);
      acc[0] += c; acc[1] += sa; acc[2] += sd; acc[3] += sd2;
    }
  }
  __shared__ double red[LOSS_THREADS / 32][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
  if ((threadIdx.x & 31) == 0)
    for (int q = 0; q < 4; ++q) red[threadIdx.x >> 5][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += red[w][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

struct LossScalars {
  float l1, silog, m1, inv_n;
};
__device__ __forceinline__ LossScalars loss_scalars(const double* sums, float lam) {
  double n = sums[0];
  LossScalars r;
  if (n <= 0.0) {  // no valid pixel: the reference's mean over an empty selection is NaN
    r.l1 = r.silog = r.m1 = __int_as_float(0x7fc00000);
    r.inv_n = 0.f;
    return r;
  }
  double m1 = sums[2] / n, m2 = sums[3] / n;
  double v = m2 - (double)lam * m1 * m1;
  r.l1 = (float)(sums[1] / n);
  r.silog = (float)sqrt(v > 0.0 ? v : 0.0);
  r.m1 = (float)m1;
  r.inv_n = (float)(1.0 / n);
  return r;
}

__global__ void loss_value_kernel(const double* __restrict__ sums, float l1_w, float silog_w, float lam,
                                  float* __restrict__ out) {
  LossScalars r = loss_scalars(sums, lam);
  float loss = l1_w * r.l1;
  if (silog_w != 0.f) loss += silog_w * r.silog;
  out[0] = loss;
  out[1] = r.l1;
  out[2] = r.silog;
}

__device__ __forceinline__ float loss_grad1(float p, float g, float scale, float eps, int use_mask,
                                            const LossScalars& r, float l1_w, float silog_w, float lam, float gs) {
  if (use_mask && g == 0.0f) return 0.f;
  float ps = p * scale, gsc = g * scale;
  float diff = ps - gsc;
  float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  float grad = l1_w * scale * sgn * r.inv_n;
  if (silog_w != 0.f && ps >= eps) {
    // d sqrt(v)/dp = (d - lam*m1) / (N * sqrt(v) * p);  v <= 0 reproduces the reference's NaN/inf
    float d = logf(ps) - logf(fmaxf(gsc, eps));
    grad += silog_w * scale * (d - lam * r.m1) * r.inv_n / (r.silog * ps);
  }
  return grad * gs;
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n, float scale,
                     float eps, int use_mask, const double* __restrict__ sums, float l1_w, float silog_w, float lam,
                     const float* __restrict__ grad_scale, float* __restrict__ dpred) {
  const LossScalars r = loss_scalars(sums, lam);
  const float gs = grad_scale ? *grad_scale : 1.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = ld4(pred + 4 * i), g = ld4(gt + 4 * i), o;
    o.x = loss_grad1(p.x, g.x, scale, eps, use_mask, r, l1_w, silog_w, lam, gs);
    o.y = loss_grad1(p.y, g.y, scale, eps, use_mask, r, l1_w, silog_w, lam, gs);
    o.z = loss_grad1(p.z, g.z, scale, eps, use_mask, r, l1_w, silog_w, lam, gs);
    o.w = loss_grad1(p.w, g.w, scale, eps, use_mask, r, l1_w, silog_w, lam, gs);
    st4(dpred + 4 * i, o);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i)
      dpred[i] = loss_grad1(pred[i], gt[i], scale, eps, use_mask, r, l1_w, silog_w, lam, gs);
}

int loss_grid(long long n) {
  long long blocks = (n / 4 + LOSS_THREADS - 1) / LOSS_THREADS;
  long long cap = (long long)adp::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int adp_depth_loss_sums(const float* pred, const float* gt, int64_t n, float scale, float eps,
                                   int use_mask, double* sums, void* stream) {
  ADP_CHECK_ARG(pred && gt && sums && n >= 0, "loss_sums: bad arguments");
  ADP_CHECK_ARG(((uintptr_t)pred % 16 == 0) && ((uintptr_t)gt % 16 == 0), "loss_sums: pointers must be 16-byte aligned");
  adp::ProfScope prof(adp::PROF_LOSS, (cudaStream_t)stream, (double)n * 8.0);            // pred, gt in
  loss_sums_kernel<<<loss_grid(n), LOSS_THREADS, 0, (cudaStream_t)stream>>>(pred, gt, n, scale, eps, use_mask, sums);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

extern "C" int adp_depth_loss_value(const double* sums, float l1_w, float silog_w, float lam, float* loss_out,
                                    void* stream) {
  ADP_CHECK_ARG(sums && loss_out, "loss_value: null pointer");
  loss_value_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, l1_w, silog_w, lam, loss_out);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

extern "C" int adp_depth_loss_backward(const float* pred, const float* gt, int64_t n, float scale, float eps,
                                       int use_mask, const double* sums, float l1_w, float silog_w, float lam,
                                       const float* grad_scale, float* dpred, void* stream) {
  ADP_CHECK_ARG(pred && gt && sums && dpred && n >= 0, "loss_backward: bad arguments");
  adp::ProfScope prof(adp::PROF_LOSS, (cudaStream_t)stream, (double)n * 4.0);            // dpred out (pred, gt counted by loss_sums)
  ADP_CHECK_ARG(((uintptr_t)pred % 16 == 0) && ((uintptr_t)gt % 16 == 0) && ((uintptr_t)dpred % 16 == 0),
                "loss_backward: pointers must be 16-byte aligned");
  loss_backward_kernel<<<loss_grid(n), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      pred, gt, n, scale, eps, use_mask, sums, l1_w, silog_w, lam, grad_scale, dpred);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}
