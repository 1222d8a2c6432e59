// Global-norm gradient clipping + AdamW as two multi-tensor kernels.
//
// Replaces torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) and torch.optim.AdamW.step()
// of the reference training loop (train.py:471-476, :689-691): one pass accumulates sum g^2 over
// every tensor, the second applies clip coefficient, decoupled weight decay, moment updates and
// the parameter update, reading p,g,m,v and writing p,m,v exactly once.
#include <math.h>
#include "adp_common.cuh"

namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_ELEMS_PER_BLOCK = OPT_THREADS * 16;
constexpr int OPT_MAX_TENSORS = 24;

struct TensorTable {
  adp_tensor_ref refs[OPT_MAX_TENSORS];
  int block_start[OPT_MAX_TENSORS + 1];
  int n;
};

__device__ __forceinline__ int find_tensor(const TensorTable& t, int block) {
  int i = 0;
  while (i + 1 < t.n && block >= t.block_start[i + 1]) ++i;
  return i;
}

__global__ void __launch_bounds__(OPT_THREADS)
grad_sumsq_kernel(const __grid_constant__ TensorTable tab, double* __restrict__ sumsq) {
  const int ti = find_tensor(tab, blockIdx.x);
  const adp_tensor_ref r = tab.refs[ti];
  const long long base = (long long)(blockIdx.x - tab.block_start[ti]) * OPT_ELEMS_PER_BLOCK;
  const long long end = min(base + (long long)OPT_ELEMS_PER_BLOCK, (long long)r.n);
  float acc = 0.f;
  if (((uintptr_t)r.g & 15) == 0) {
    long long i = base + threadIdx.x * 4;
    for (; i + 3 < end; i += OPT_THREADS * 4) {
      float4 g = ld4(r.g + i);
      acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
    for (; i < end; ++i) acc += r.g[i] * r.g[i];  // at most one thread has a 1-3 element tail
  } else {
    for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) acc += r.g[i] * r.g[i];
  }
  double d = warp_sum((double)acc);
  __shared__ double red[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
    atomicAdd(sumsq, t);
  }
}

struct AdamArgs {
  float max_norm, lr, beta1, beta2, eps, decay_mul, step_size, inv_bc2_sqrt;
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float coef, const AdamArgs& a) {
  g *= coef;
  p *= a.decay_mul;
  m = m + (g - m) * (1.f - a.beta1);
  v = v * a.beta2 + (1.f - a.beta2) * g * g;
  float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

// Device-resident step counter for CUDA-graph replay: scalars = {step_size, 1/sqrt(bias_correction2)}
__global__ void adam_tick_kernel(int* __restrict__ step, float* __restrict__ scalars, float lr, float beta1, float beta2) {
  const int t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const double bc2 = 1.0 - pow((double)beta2, (double)t);
  scalars[0] = (float)((double)lr / bc1);
  scalars[1] = (float)(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(OPT_THREADS)
clip_adamw_kernel(const __grid_constant__ TensorTable tab, const double* __restrict__ sumsq, AdamArgs a,
                  float* __restrict__ norm_out, const float* __restrict__ dev_scalars) {
  if (dev_scalars) { a.step_size = dev_scalars[0]; a.inv_bc2_sqrt = dev_scalars[1]; }
  const float total = (float)sqrt(*sumsq);
  float coef = a.max_norm > 0.f ? a.max_norm / (total + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f);
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
  const int ti = find_tensor(tab, blockIdx.x);
  const adp_tensor_ref r = tab.refs[ti];
  const long long base = (long long)(blockIdx.x - tab.block_start[ti]) * OPT_ELEMS_PER_BLOCK;
  const long long end = min(base + (long long)OPT_ELEMS_PER_BLOCK, (long long)r.n);
  __nv_bfloat16* p16 = reinterpret_cast<__nv_bfloat16*>(r.p_bf16);
  const bool aligned = ((((uintptr_t)r.p) | ((uintptr_t)r.g) | ((uintptr_t)r.m) | ((uintptr_t)r.v)) & 15) == 0 &&
                       (((uintptr_t)r.p_bf16) & 7) == 0;
  if (aligned) {
    long long i = base + threadIdx.x * 4;
    for (; i + 3 < end; i += OPT_THREADS * 4) {
      float4 p = ld4(r.p + i), g = ld4(r.g + i), m = ld4(r.m + i), v = ld4(r.v + i);
      adam1(p.x, g.x, m.x, v.x, coef, a);
      adam1(p.y, g.y, m.y, v.y, coef, a);
      adam1(p.z, g.z, m.z, v.z, coef, a);
      adam1(p.w, g.w, m.w, v.w, coef, a);
      st4(r.p + i, p); st4(r.m + i, m); st4(r.v + i, v);
      if (p16) {   // i % 4 == 0 and the mirror is 8-byte aligned with p 16-byte aligned
        uint2 u;
        u.x = pack_bf16x2_(p.x, p.y);
        u.y = pack_bf16x2_(p.z, p.w);
        *reinterpret_cast<uint2*>(p16 + i) = u;
      }
    }
    for (; i < end; ++i) {
      adam1(r.p[i], r.g[i], r.m[i], r.v[i], coef, a);
      if (p16) p16[i] = __float2bfloat16(r.p[i]);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) {
      adam1(r.p[i], r.g[i], r.m[i], r.v[i], coef, a);
      if (p16) p16[i] = __float2bfloat16(r.p[i]);
    }
  }
}

template <class F>
int for_each_table(const adp_tensor_ref* refs, int n_tensors, F&& launch) {
  int i = 0;
  while (i < n_tensors) {
    TensorTable tab;
    tab.n = 0;
    int blocks = 0;
    while (i < n_tensors && tab.n < OPT_MAX_TENSORS) {
      if (refs[i].n > 0) {
        tab.refs[tab.n] = refs[i];
        tab.block_start[tab.n] = blocks;
        blocks += (int)((refs[i].n + OPT_ELEMS_PER_BLOCK - 1) / OPT_ELEMS_PER_BLOCK);
        ++tab.n;
      }
      ++i;
    }
    tab.block_start[tab.n] = blocks;
    if (tab.n > 0) ADP_TRY(launch(tab, blocks));
  }
  return ADP_OK;
}

}  // namespace

extern "C" int adp_grad_sumsq(const adp_tensor_ref* refs_host, int n_tensors, double* sumsq, void* stream) {
  ADP_CHECK_ARG(refs_host && n_tensors > 0 && sumsq, "grad_sumsq: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  double elems = 0.0;
  for (int i = 0; i < n_tensors; ++i) elems += (double)refs_host[i].n;
  adp::ProfScope prof(adp::PROF_OPTIM, s, elems * 4.0);                                  // g in
  return for_each_table(refs_host, n_tensors, [&](const TensorTable& tab, int blocks) -> int {
    grad_sumsq_kernel<<<blocks, OPT_THREADS, 0, s>>>(tab, sumsq);
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  });
}

extern "C" int adp_clip_adamw_step(const adp_tensor_ref* refs_host, int n_tensors, const double* sumsq,
                                   float max_norm, float lr, float beta1, float beta2, float eps,
                                   float weight_decay, int step, float* norm_out, void* stream) {
  ADP_CHECK_ARG(refs_host && n_tensors > 0 && sumsq && step >= 1, "clip_adamw_step: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  double elems = 0.0;
  for (int i = 0; i < n_tensors; ++i) elems += (double)refs_host[i].n;
  adp::ProfScope prof(adp::PROF_OPTIM, s, elems * 4.0 * 7.0);                            // p, g, m, v in; p, m, v out
  AdamArgs a;
  a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.decay_mul = 1.f - lr * weight_decay;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  return for_each_table(refs_host, n_tensors, [&](const TensorTable& tab, int blocks) -> int {
    clip_adamw_kernel<<<blocks, OPT_THREADS, 0, s>>>(tab, sumsq, a, norm_out, nullptr);
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  });
}

// Same update with the step counter kept on the device (step_dev: int, incremented here; scratch: 2 floats), so
// that the call can be captured once in a CUDA graph and replayed.
extern "C" int adp_clip_adamw_step_graph(const adp_tensor_ref* refs_host, int n_tensors, const double* sumsq,
                                         float max_norm, float lr, float beta1, float beta2, float eps,
                                         float weight_decay, int* step_dev, float* scratch, float* norm_out,
                                         void* stream) {
  ADP_CHECK_ARG(refs_host && n_tensors > 0 && sumsq && step_dev && scratch, "clip_adamw_step_graph: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  double elems = 0.0;
  for (int i = 0; i < n_tensors; ++i) elems += (double)refs_host[i].n;
  adp::ProfScope prof(adp::PROF_OPTIM, s, elems * 4.0 * 7.0);                            // p, g, m, v in; p, m, v out
  AdamArgs a;
  a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.decay_mul = 1.f - lr * weight_decay;
  a.step_size = 0.f; a.inv_bc2_sqrt = 0.f;
  adam_tick_kernel<<<1, 1, 0, s>>>(step_dev, scratch, lr, beta1, beta2);
  ADP_LAUNCH_CHECK();
  return for_each_table(refs_host, n_tensors, [&](const TensorTable& tab, int blocks) -> int {
    clip_adamw_kernel<<<blocks, OPT_THREADS, 0, s>>>(tab, sumsq, a, norm_out, scratch);
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  });
}
