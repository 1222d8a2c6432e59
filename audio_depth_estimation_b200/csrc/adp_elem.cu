// Bandwidth-bound elementwise / reduction kernels of the U-Net: BatchNorm statistics and
// normalise+activation, their backward, the output head, weight casts.
//
// Replaces, for models/unetbaseline_model.py:187-229: nn.BatchNorm2d (:190,:192,:219),
// nn.LeakyReLU(0.2, inplace) (:189), nn.ReLU(inplace) (:191), torch.cat (:235, eliminated:
// the decoder reads the two halves of the concat as two tensors) and nn.Sigmoid/ReLU head
// (:201-206).  All activation tensors are NHWC, viewed as [rows = B*H*W, C].
#include <stdlib.h>
#include "adp_common.cuh"

namespace {

constexpr int EW_THREADS = 256;

// ------------------------------------------------------------------ BN statistics
// block (32, 8): threadIdx.x owns 4 consecutive channels, rows strided over threadIdx.y / blockIdx.y
template <class T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ x, long long rows, int C, double* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (c < C) {
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      float4 v = ld4(x + r * C + c);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]);
      q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  }
  __shared__ float red[8][32][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[threadIdx.y][threadIdx.x][i] = s[i];
    red[threadIdx.y][threadIdx.x][4 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double a = 0.0, b = 0.0;
      for (int y = 0; y < 8; ++y) {
        a += (double)red[y][threadIdx.x][i];
        b += (double)red[y][threadIdx.x][4 + i];
      }
      atomicAdd(&sums[c + i], a);
      atomicAdd(&sums[C + c + i], b);
    }
  }
}

// Per-channel BatchNorm coefficients from the accumulated sums (training) or the running statistics (eval).  One
// definition, used by the stand-alone finalize kernel AND by every thread of the fused normalise kernel, so that the
// stored scale/shift (read by the backward pass) are bit-identical with what the forward pass applied.
struct BnCoef { float mean, invstd, scale, shift, unbiased; };
__device__ __forceinline__ BnCoef bn_coef(const adp::BnFin& f, int C, int c) {
  BnCoef o;
  if (f.training) {
    const double m = f.sums[c] * f.inv_rows;                 // (no double division on the per-channel path)
    double var = fma(-m, m, f.sums[C + c] * f.inv_rows);
    if (var < 0.0) var = 0.0;
    o.mean = (float)m;
    o.invstd = 1.f / sqrtf((float)var + f.eps);
    o.unbiased = (float)var * f.unbias;
  } else {
    o.mean = f.rm[c] - (f.mean_offset ? f.mean_offset[c] : 0.f);     // statistics of the stored (offset) tensor
    o.invstd = 1.f / sqrtf(f.rv[c] + f.eps);
    o.unbiased = 0.f;
  }
  o.scale = f.gamma[c] * o.invstd;
  o.shift = f.beta[c] - o.mean * o.scale;
  return o;
}
__device__ __forceinline__ void bn_store(const adp::BnFin& f, int c, const BnCoef& o) {
  if (f.training && f.rm) {
    f.rm[c] = (1.f - f.momentum) * f.rm[c] + f.momentum * (o.mean + (f.mean_offset ? f.mean_offset[c] : 0.f));
    f.rv[c] = (1.f - f.momentum) * f.rv[c] + f.momentum * o.unbiased;
  }
  f.scale[c] = o.scale;
  f.shift[c] = o.shift;
  f.mean[c] = o.mean;
  f.invstd[c] = o.invstd;
}

// (fallback for channel counts the fused 8-wide kernels do not cover: 2048 % C != 0)
__global__ void bn_finalize_kernel(const adp::BnFin f, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_store(f, c, bn_coef(f, C, c));
}

// ------------------------------------------------------------------ normalise + activation
template <class T>
__global__ void __launch_bounds__(EW_THREADS)
affine_act_kernel(const T* __restrict__ x, long long n4, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, float slope0, T* __restrict__ out0, float slope1,
                  T* __restrict__ out1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ld4(x + 4 * i);
    if (scale) {
      int c = (int)((4 * i) % C);
      float4 sc = ld4(scale + c), sh = ld4(shift + c);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
      v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    st4(out0 + 4 * i, make_float4(lrelu(v.x, slope0), lrelu(v.y, slope0), lrelu(v.z, slope0), lrelu(v.w, slope0)));
    if (out1)
      st4(out1 + 4 * i, make_float4(lrelu(v.x, slope1), lrelu(v.y, slope1), lrelu(v.z, slope1), lrelu(v.w, slope1)));
  }
}

// ------------------------------------------------------------------ backward: activation (+BN)
template <class T>
__device__ __forceinline__ float4 load_gz(const T* x, long long off, int c, const float* scale, const float* shift,
                                          const T* gA, float slope0, const T* gB, float slope1, float4& xv) {
  xv = ld4(x + off);
  float4 z = xv;
  if (scale) {
    float4 sc = ld4(scale + c), sh = ld4(shift + c);
    z.x = fmaf(z.x, sc.x, sh.x); z.y = fmaf(z.y, sc.y, sh.y);
    z.z = fmaf(z.z, sc.z, sh.z); z.w = fmaf(z.w, sc.w, sh.w);
  }
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gA) {
    float4 a = ld4(gA + off);
    g.x = a.x * lrelu_grad(z.x, slope0); g.y = a.y * lrelu_grad(z.y, slope0);
    g.z = a.z * lrelu_grad(z.z, slope0); g.w = a.w * lrelu_grad(z.w, slope0);
  }
  if (gB) {
    float4 b = ld4(gB + off);
    g.x = fmaf(b.x, lrelu_grad(z.x, slope1), g.x); g.y = fmaf(b.y, lrelu_grad(z.y, slope1), g.y);
    g.z = fmaf(b.z, lrelu_grad(z.z, slope1), g.z); g.w = fmaf(b.w, lrelu_grad(z.w, slope1), g.w);
  }
  return g;
}

template <class T>
__global__ void __launch_bounds__(256)
act_bn_bwd_reduce_kernel(const T* __restrict__ x, long long rows, int C, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const T* __restrict__ gA, float slope0,
                         const T* __restrict__ gB, float slope1, double* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (c < C) {
    float4 mu = ld4(mean + c), is = ld4(invstd + c);
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
      float4 xv;
      float4 g = load_gz(x, r * C + c, c, scale, shift, gA, slope0, gB, slope1, xv);
      s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
      q[0] = fmaf(g.x, (xv.x - mu.x) * is.x, q[0]); q[1] = fmaf(g.y, (xv.y - mu.y) * is.y, q[1]);
      q[2] = fmaf(g.z, (xv.z - mu.z) * is.z, q[2]); q[3] = fmaf(g.w, (xv.w - mu.w) * is.w, q[3]);
    }
  }
  __shared__ float red[8][32][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[threadIdx.y][threadIdx.x][i] = s[i];
    red[threadIdx.y][threadIdx.x][4 + i] = q[i];
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double a = 0.0, b = 0.0;
      for (int y = 0; y < 8; ++y) {
        a += (double)red[y][threadIdx.x][i];
        b += (double)red[y][threadIdx.x][4 + i];
      }
      atomicAdd(&sums[c + i], a);
      atomicAdd(&sums[C + c + i], b);
    }
  }
}

template <class T>
__global__ void __launch_bounds__(EW_THREADS)
act_bn_bwd_apply_kernel(const T* __restrict__ x, long long n4, long long rows, int C,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const T* __restrict__ gA, float slope0, const T* __restrict__ gB, float slope1,
                        const double* __restrict__ sums, int mode, T* __restrict__ dx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv_m = 1.f / (float)rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = (int)((4 * i) % C);
    float4 xv;
    float4 g = load_gz(x, 4 * i, c, scale, shift, gA, slope0, gB, slope1, xv);
    if (mode == 1) {
      float4 sc = ld4(scale + c);
      g.x *= sc.x; g.y *= sc.y; g.z *= sc.z; g.w *= sc.w;
    } else if (mode == 2) {
      float4 sc = ld4(scale + c), mu = ld4(mean + c), is = ld4(invstd + c);
      float s1[4], s2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s1[k] = (float)sums[c + k] * inv_m;
        s2[k] = (float)sums[C + c + k] * inv_m;
      }
      g.x = sc.x * (g.x - s1[0] - (xv.x - mu.x) * is.x * s2[0]);
      g.y = sc.y * (g.y - s1[1] - (xv.y - mu.y) * is.y * s2[1]);
      g.z = sc.z * (g.z - s1[2] - (xv.z - mu.z) * is.z * s2[2]);
      g.w = sc.w * (g.w - s1[3] - (xv.w - mu.w) * is.w * s2[3]);
    }
    st4(dx + 4 * i, g);
  }
}


// ================================================================== 8-wide variants (C % 8 == 0)
// Column reductions use blockDim = (tx, 256/tx) with tx = min(32, C/8): every thread owns 8 consecutive
// channels and strides over rows, so narrow tensors (C = 64) still use all 256 threads.
// Final stage: shuffle over the rows held by one warp, then one (channel, value) per thread.
__device__ __forceinline__ void col_reduce16(float (&v)[16], int C, int c0_block, double* __restrict__ sums) {
  __shared__ float red[8][32][17];
  const int tx = blockDim.x;
  const int tid = threadIdx.y * tx + threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int off = tx; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  if (lane < tx) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp][lane][i] = v[i];
  }
  __syncthreads();
  // columns per warp-row: tx >= 32 means every lane is a distinct column group
  const int groups = tx < 32 ? tx : 32;
  const int nwarps = (tx * blockDim.y) >> 5;
  for (int item = tid; item < groups * 16; item += tx * blockDim.y) {
    const int x = item >> 4, i = item & 15;
    double a = 0.0;
    if (tx < 32) {
      for (int w = 0; w < nwarps; ++w) a += (double)red[w][x][i];
    } else {
      for (int w = 0; w < nwarps; ++w) a += (double)red[w][x][i];
    }
    const int c = c0_block + x * 8;
    if (c < C) atomicAdd(&sums[(i < 8 ? 0 : C) + c + (i & 7)], a);
  }
}

template <class T>
__global__ void __launch_bounds__(256)
bn_stats8_kernel(const T* __restrict__ x, long long rows, int C, double* __restrict__ sums) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  if (c < C) {
    const long long step = (long long)gridDim.y * blockDim.y;
    long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    auto add = [&](const float8& a) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] += a.v[i]; v[8 + i] = fmaf(a.v[i], a.v[i], v[8 + i]); }
    };
    for (; r + 3 * step < rows; r += 4 * step) {   // four independent 16-byte loads in flight
      typename Raw8<T>::type q0 = ldraw8(x + r * C + c), q1 = ldraw8(x + (r + step) * C + c),
                             q2 = ldraw8(x + (r + 2 * step) * C + c), q3 = ldraw8(x + (r + 3 * step) * C + c);
      add(cvt8(q0)); add(cvt8(q1)); add(cvt8(q2)); add(cvt8(q3));
    }
    for (; r < rows; r += step) add(ld8(x + r * C + c));
  }
  col_reduce16(v, C, blockIdx.x * blockDim.x * 8, sums);
}

// Streaming kernels: a block covers 2048 consecutive elements per iteration, so when 2048 % C == 0 a
// thread meets the SAME 8 channels in every iteration and its per-channel coefficients stay in registers.
template <class T, bool FIXED>
__global__ void __launch_bounds__(EW_THREADS)
affine_act8_kernel(const T* __restrict__ x, long long n8, int C, const float* __restrict__ scale,
                   const float* __restrict__ shift, float slope0, T* __restrict__ out0, float slope1, T* __restrict__ out1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float8 sc, sh;
  if (FIXED && scale) { const int c = (threadIdx.x * 8) % C; sc = ld8(scale + c); sh = ld8(shift + c); }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float8 v = ld8(x + 8 * i);
    if (scale) {
      if (!FIXED) { const int c = (int)((8 * i) % C); sc = ld8(scale + c); sh = ld8(shift + c); }
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = fmaf(v.v[k], sc.v[k], sh.v[k]);
    }
    float8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = lrelu(v.v[k], slope0);
    st8(out0 + 8 * i, o);
    if (out1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = lrelu(v.v[k], slope1);
      st8(out1 + 8 * i, o);
    }
  }
}

// BatchNorm finalize + normalise + activation in one launch (2048 % C == 0): every thread derives the coefficients of
// its own 8 channels from the accumulated sums, block 0 additionally stores scale/shift/mean/invstd for the backward
// pass and updates the running statistics.
template <class T>
__global__ void __launch_bounds__(EW_THREADS)
bn_affine_act8_kernel(const T* __restrict__ x, long long n8, int C, const adp::BnFin f, float slope0, T* __restrict__ out0,
                      float slope1, T* __restrict__ out1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  // the block derives the C coefficient pairs once (C <= 2048), every thread then picks up its 8 channels
  __shared__ float2 coef_s[2048];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const BnCoef o = bn_coef(f, C, c);
    coef_s[c] = make_float2(o.scale, o.shift);
    if (blockIdx.x == 0) bn_store(f, c, o);      // (training: nobody reads rm/rv; eval: nobody writes them)
  }
  __syncthreads();
  float8 sc, sh;
  {
    const int c = (threadIdx.x * 8) % C;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float2 t = coef_s[c + k]; sc.v[k] = t.x; sh.v[k] = t.y; }
  }
  // (4 independent loads in flight per thread: the mid-size layers, 16-67 MB, are latency- not bandwidth-bound)
  constexpr int U = 4;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += U * stride) {
    typename Raw8<T>::type raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < n8) raw[u] = ldraw8(x + 8 * (i0 + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n8) break;
      float8 v = cvt8(raw[u]);
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = fmaf(v.v[k], sc.v[k], sh.v[k]);
      float8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = lrelu(v.v[k], slope0);
      st8(out0 + 8 * i, o);
      if (out1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = lrelu(v.v[k], slope1);
        st8(out1 + 8 * i, o);
      }
    }
  }
}

struct GzIn { float8 x, a, b; };
template <class T>
__device__ __forceinline__ GzIn gz_load(const T* x, const T* gA, const T* gB, long long off) {
  GzIn r;
  r.x = ld8(x + off);
  if (gA) r.a = ld8(gA + off);
  if (gB) r.b = ld8(gB + off);
  return r;
}
__device__ __forceinline__ float8 gz_compute(const GzIn& in, const float8* sc, const float8* sh, bool hasA, float slope0,
                                             bool hasB, float slope1) {
  float8 z = in.x;
  if (sc) {
#pragma unroll
    for (int k = 0; k < 8; ++k) z.v[k] = fmaf(z.v[k], sc->v[k], sh->v[k]);
  }
  float8 g;
#pragma unroll
  for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
  if (hasA) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = in.a.v[k] * lrelu_grad(z.v[k], slope0);
  }
  if (hasB) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = fmaf(in.b.v[k], lrelu_grad(z.v[k], slope1), g.v[k]);
  }
  return g;
}

// Register-free prefetch: every thread owns private 16-byte shared-memory slots that it fills with cp.async and
// later reads back itself (no block synchronisation), so RED_ROWS rows x up to 3 tensors are in flight per thread.
constexpr int RED_ROWS = 3, RED_STAGES = 2;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class T>
__global__ void __launch_bounds__(256)
act_bn_bwd_reduce8_kernel(const T* __restrict__ x, long long rows, int C, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ invstd, const T* __restrict__ gA, float slope0,
                          const T* __restrict__ gB, float slope1, double* __restrict__ sums) {
  constexpr int CH16 = (int)(sizeof(T) * 8 / 16);                 // 16-byte chunks per 8 elements
  extern __shared__ __align__(16) unsigned char red_smem[];      // [STAGES][ROWS][3][CH16][256] x 16 B
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  auto slot = [&](int stage, int j, int t, int h) -> unsigned char* {
    return red_smem + ((((size_t)(stage * RED_ROWS + j) * 3 + t) * CH16 + h) * 256 + tid) * 16;
  };
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  if (c < C) {
    float8 sc, sh;
    if (scale) { sc = ld8(scale + c); sh = ld8(shift + c); }
    const float8 mu = ld8(mean + c);
    const float8* scp = scale ? &sc : nullptr;
    const long long step = (long long)gridDim.y * blockDim.y;
    auto issue = [&](int stage, long long r0) {
#pragma unroll
      for (int j = 0; j < RED_ROWS; ++j) {
        const long long r = r0 + j * step;
        if (r < rows) {
          const long long off = r * C + c;
#pragma unroll
          for (int h = 0; h < CH16; ++h) {
            cp_async16(slot(stage, j, 0, h), reinterpret_cast<const unsigned char*>(x + off) + 16 * h);
            if (gA) cp_async16(slot(stage, j, 1, h), reinterpret_cast<const unsigned char*>(gA + off) + 16 * h);
            if (gB) cp_async16(slot(stage, j, 2, h), reinterpret_cast<const unsigned char*>(gB + off) + 16 * h);
          }
        }
      }
      cp_async_commit();
    };
    auto read8 = [&](int stage, int j, int t) -> float8 {
      if (CH16 == 1) return cvt8(*reinterpret_cast<const uint4*>(slot(stage, j, t, 0)));
      raw8f q;
      q.a = *reinterpret_cast<const float4*>(slot(stage, j, t, 0));
      q.b = *reinterpret_cast<const float4*>(slot(stage, j, t, CH16 - 1));
      return cvt8(q);
    };
    long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    int stage = 0;
    issue(0, r);
    while (r < rows) {
      const long long rn = r + RED_ROWS * step;
      issue(stage ^ 1, rn);                 // empty group when rn >= rows
      cp_async_wait<1>();
#pragma unroll
      for (int j = 0; j < RED_ROWS; ++j) {
        if (r + j * step < rows) {
          GzIn in;
          in.x = read8(stage, j, 0);
          if (gA) in.a = read8(stage, j, 1);
          if (gB) in.b = read8(stage, j, 2);
          const float8 g = gz_compute(in, scp, &sh, gA != nullptr, slope0, gB != nullptr, slope1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] += g.v[i];
            v[8 + i] = fmaf(g.v[i], in.x.v[i] - mu.v[i], v[8 + i]);
          }
        }
      }
      r = rn;
      stage ^= 1;
    }
    cp_async_wait<0>();
    const float8 is = ld8(invstd + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 + i] *= is.v[i];
  }
  col_reduce16(v, C, blockIdx.x * blockDim.x * 8, sums);
}

// dx = A*gz + Bc*x + Cc  with  mode 0: A=1,Bc=Cc=0;  1: A=scale;  2: A=scale, Bc=-scale*invstd*s2/M,
// Cc = -scale*s1/M + scale*invstd*mean*s2/M   (the batch-statistics BatchNorm backward)
template <class T, bool FIXED>
__global__ void __launch_bounds__(EW_THREADS)
act_bn_bwd_apply8_kernel(const T* __restrict__ x, long long n8, long long rows, int C, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const T* __restrict__ gA, float slope0,
                         const T* __restrict__ gB, float slope1, const double* __restrict__ sums, int mode,
                         T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv_m = 1.f / (float)rows;
  if (dgamma && blockIdx.x == 0) {     // dbeta = sum gz, dgamma = sum gz * xhat
    for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta[c] = (float)sums[c]; dgamma[c] = (float)sums[C + c]; }
  }
  float8 sc, sh, cA, cB, cC;
  auto coeffs = [&](int c) {
    if (scale) { sc = ld8(scale + c); sh = ld8(shift + c); }
#pragma unroll
    for (int k = 0; k < 8; ++k) { cA.v[k] = 1.f; cB.v[k] = 0.f; cC.v[k] = 0.f; }
    if (mode >= 1) cA = sc;
    if (mode == 2) {
      const float8 mu = ld8(mean + c), is = ld8(invstd + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float s1 = (float)sums[c + k] * inv_m, s2 = (float)sums[C + c + k] * inv_m;
        cB.v[k] = -sc.v[k] * is.v[k] * s2;
        cC.v[k] = -sc.v[k] * s1 - cB.v[k] * mu.v[k];
      }
    }
  };
  if (FIXED) coeffs((threadIdx.x * 8) % C);
  const float8* scp = scale ? &sc : nullptr;
  auto finish = [&](const GzIn& in, long long i) {
    float8 g = gz_compute(in, scp, &sh, gA != nullptr, slope0, gB != nullptr, slope1);
#pragma unroll
    for (int k = 0; k < 8; ++k) g.v[k] = fmaf(cA.v[k], g.v[k], fmaf(cB.v[k], in.x.v[k], cC.v[k]));
    st8(dx + 8 * i, g);
  };
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (FIXED) {
    // same register-free cp.async prefetch as the reduction: RED_ROWS iterations x up to 3 tensors in flight
    constexpr int CH16 = (int)(sizeof(T) * 8 / 16);
    extern __shared__ __align__(16) unsigned char red_smem[];
    const int tid = threadIdx.x;
    auto slot = [&](int stage, int j, int t, int h) -> unsigned char* {
      return red_smem + ((((size_t)(stage * RED_ROWS + j) * 3 + t) * CH16 + h) * 256 + tid) * 16;
    };
    auto issue = [&](int stage, long long i0) {
#pragma unroll
      for (int j = 0; j < RED_ROWS; ++j) {
        const long long ii = i0 + j * stride;
        if (ii < n8) {
#pragma unroll
          for (int h = 0; h < CH16; ++h) {
            cp_async16(slot(stage, j, 0, h), reinterpret_cast<const unsigned char*>(x + 8 * ii) + 16 * h);
            if (gA) cp_async16(slot(stage, j, 1, h), reinterpret_cast<const unsigned char*>(gA + 8 * ii) + 16 * h);
            if (gB) cp_async16(slot(stage, j, 2, h), reinterpret_cast<const unsigned char*>(gB + 8 * ii) + 16 * h);
          }
        }
      }
      cp_async_commit();
    };
    auto read8 = [&](int stage, int j, int t) -> float8 {
      if (CH16 == 1) return cvt8(*reinterpret_cast<const uint4*>(slot(stage, j, t, 0)));
      raw8f q;
      q.a = *reinterpret_cast<const float4*>(slot(stage, j, t, 0));
      q.b = *reinterpret_cast<const float4*>(slot(stage, j, t, CH16 - 1));
      return cvt8(q);
    };
    int stage = 0;
    issue(0, i);
    while (i < n8) {
      const long long in_ = i + RED_ROWS * stride;
      issue(stage ^ 1, in_);
      cp_async_wait<1>();
#pragma unroll
      for (int j = 0; j < RED_ROWS; ++j) {
        const long long ii = i + j * stride;
        if (ii < n8) {
          GzIn q;
          q.x = read8(stage, j, 0);
          if (gA) q.a = read8(stage, j, 1);
          if (gB) q.b = read8(stage, j, 2);
          finish(q, ii);
        }
      }
      i = in_;
      stage ^= 1;
    }
    cp_async_wait<0>();
    return;
  }
  for (; i < n8; i += stride) {
    if (!FIXED) coeffs((int)((8 * i) % C));
    finish(gz_load(x, gA, gB, 8 * i), i);
  }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    dbeta[c] = (float)sums[c];
    dgamma[c] = (float)sums[C + c];
  }
}


// ================================================================== small tensors: one launch per BatchNorm layer
// The deep levels (rows = B*H*W <= a few thousand) are launch-bound: statistics + normalise (forward) and reduce + apply
// (backward) cost 5-8 us per launch for a few hundred KB.  Here one block owns an 8-channel slab of the whole tensor,
// stages it in shared memory (16 bytes per row and tensor), reduces over the rows and applies -- each tensor is read
// from L2 / HBM exactly once and the layer is ONE launch.
constexpr int SMALL_THREADS = 256;
constexpr int SMALL_U = 4;            // rows a thread has in flight (rows <= 1024: the whole first pass)

template <class T>
__device__ __forceinline__ void slab_block_reduce16(float (&v)[16], double (&tot)[16]) {
  __shared__ float red[SMALL_THREADS / 32][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();                                    // (red may still be read by a previous call)
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double a = 0.0;
    for (int w = 0; w < SMALL_THREADS / 32; ++w) a += (double)red[w][i];
    tot[i] = a;
  }
}

// Split-K partial sums handed over un-finished by a tensor-core convolution (ConvExtras::deferred): 8 fp32 sums are read,
// cleared (the scratch stays all zero for the next split launch) and rounded to T exactly as finish_partial_kernel would
// have stored them -- the bf16 tensor in between is never written unless somebody else reads it.
__device__ __forceinline__ uint4 round_raw8(const float8& a, const bf16*) {
  uint4 u;
  u.x = pack_bf16x2_(a.v[0], a.v[1]); u.y = pack_bf16x2_(a.v[2], a.v[3]);
  u.z = pack_bf16x2_(a.v[4], a.v[5]); u.w = pack_bf16x2_(a.v[6], a.v[7]);
  return u;
}
__device__ __forceinline__ raw8f round_raw8(const float8& a, const float*) {
  raw8f r;
  r.a = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  r.b = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
  return r;
}
__device__ __forceinline__ void straw8(bf16* p, const uint4& q) { *reinterpret_cast<uint4*>(p) = q; }
__device__ __forceinline__ void straw8(float* p, const raw8f& q) {
  *reinterpret_cast<float4*>(p) = q.a;
  *reinterpret_cast<float4*>(p + 4) = q.b;
}
__device__ __forceinline__ float8 float8_zero() {
  float8 z;
#pragma unroll
  for (int i = 0; i < 8; ++i) z.v[i] = 0.f;
  return z;
}

// The innermost level (no BatchNorm) consuming un-finished split-K sums, elementwise over n8 groups of 8 channels:
//   mask == NULL: raw = round(sums), act = lrelu(raw, slope)                       (forward: e[D-1] and r[D-1])
//   mask != NULL: act = round(sums) * lrelu'(mask, slope); raw is not written     (backward: dL/de from dL/dr and e)
template <class T>
__global__ void __launch_bounds__(EW_THREADS)
finish_act8_kernel(float* __restrict__ partial, long long n8, const T* __restrict__ mask, float slope, T* __restrict__ raw,
                   T* __restrict__ act) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float8 a = ld8(partial + 8 * i);
    st8(partial + 8 * i, float8_zero());
    const typename Raw8<T>::type q = round_raw8(a, (const T*)nullptr);
    const float8 v = cvt8(q);
    float8 o;
    if (mask) {
      const float8 m = ld8(mask + 8 * i);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = v.v[k] * lrelu_grad(m.v[k], slope);
    } else {
      straw8(raw + 8 * i, q);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = lrelu(v.v[k], slope);
    }
    st8(act + 8 * i, o);
  }
}

// forward: batch statistics (training) + normalise + one or two activations; slab[rows] of 8 channels in shared memory
// (partial != NULL: x is still the fp32 split-K sums [rows][C] of the convolution; it is finished here and written to xw)
template <class T>
__global__ void __launch_bounds__(SMALL_THREADS)
bn_small_fwd_kernel(const T* __restrict__ x, int rows, int C, const adp::BnFin f, float slope0, T* __restrict__ out0,
                    float slope1, T* __restrict__ out1, float* __restrict__ partial, T* __restrict__ xw) {
  extern __shared__ __align__(16) unsigned char slab_raw[];
  typedef typename Raw8<T>::type R8;
  R8* slab = reinterpret_cast<R8*>(slab_raw);
  const int c0 = blockIdx.x * 8;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  // (SMALL_U rows per thread at a time: all their loads are in flight before the first store -- the rows of a thread are
  // otherwise one dependent L2 round trip each, and the kernel is nothing but latency)
  for (int rb = threadIdx.x; rb < rows; rb += SMALL_U * SMALL_THREADS) {
    R8 q[SMALL_U];
    float8 pa[SMALL_U];
#pragma unroll
    for (int u = 0; u < SMALL_U; ++u) {
      const int r = rb + u * SMALL_THREADS;
      if (r < rows) {
        if (partial) pa[u] = ld8(partial + (size_t)r * C + c0);
        else q[u] = ldraw8(x + (size_t)r * C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < SMALL_U; ++u) {
      const int r = rb + u * SMALL_THREADS;
      if (r < rows) {
        if (partial) {
          q[u] = round_raw8(pa[u], (const T*)nullptr);
          st8(partial + (size_t)r * C + c0, float8_zero());
          straw8(xw + (size_t)r * C + c0, q[u]);       // the raw convolution output: the backward pass normalises it again
        }
        slab[r] = q[u];
        const float8 a = cvt8(q[u]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] += a.v[i]; v[8 + i] = fmaf(a.v[i], a.v[i], v[8 + i]); }
      }
    }
  }
  double tot[16];
  slab_block_reduce16<T>(v, tot);
  __shared__ float sc_s[8], sh_s[8];
  if (threadIdx.x < 8) {
    const int c = c0 + threadIdx.x;
    BnCoef o;
    const double m = tot[threadIdx.x] * f.inv_rows;
    double var = fma(-m, m, tot[8 + threadIdx.x] * f.inv_rows);
    if (var < 0.0) var = 0.0;
    o.mean = (float)m;
    o.invstd = 1.f / sqrtf((float)var + f.eps);
    o.unbiased = (float)var * f.unbias;
    o.scale = f.gamma[c] * o.invstd;
    o.shift = f.beta[c] - o.mean * o.scale;
    bn_store(f, c, o);
    sc_s[threadIdx.x] = o.scale;
    sh_s[threadIdx.x] = o.shift;
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = sc_s[i]; sh[i] = sh_s[i]; }
  for (int r = threadIdx.x; r < rows; r += SMALL_THREADS) {
    const float8 a = cvt8(slab[r]);
    float8 z, o;
#pragma unroll
    for (int i = 0; i < 8; ++i) z.v[i] = fmaf(a.v[i], sc[i], sh[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = lrelu(z.v[i], slope0);
    st8(out0 + (size_t)r * C + c0, o);
    if (out1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = lrelu(z.v[i], slope1);
      st8(out1 + (size_t)r * C + c0, o);
    }
  }
}

// backward of activation(s) + BatchNorm (mode 2: batch statistics, 1: eval): x, gA, gB staged once; dgamma / dbeta written
// pp.partial != NULL: gA is still the fp32 split-K sums of the data-gradient convolution, [rows][pp.ld] at column pp.off
// (finished here, in registers: the bf16 gA tensor has no other reader); columns [0, pp.side_c) of the same sums belong to
// another tensor (the skip-connection gradient of a decoder layer) and are finished into pp.side [rows][pp.side_c].
template <class T>
__global__ void __launch_bounds__(SMALL_THREADS)
bn_small_bwd_kernel(const T* __restrict__ x, int rows, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const T* __restrict__ gA, float slope0,
                    const T* __restrict__ gB, float slope1, int mode, T* __restrict__ dx, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, double* __restrict__ sums, const adp::BnSmallPartial pp) {
  extern __shared__ __align__(16) unsigned char slab_raw[];
  typedef typename Raw8<T>::type R8;
  R8* sx = reinterpret_cast<R8*>(slab_raw);
  R8* sg = sx + rows;                                  // gz (computed once), stored as fp32 pairs would double the slab:
  const int c0 = blockIdx.x * 8;                       // keep gA and gB raw instead and recompute gz in the second pass
  R8* sb = sg + rows;
  const float8 sc = ld8(scale + c0), sh = ld8(shift + c0), mu = ld8(mean + c0), is = ld8(invstd + c0);
  const bool hasA = gA != nullptr || pp.partial != nullptr;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  if (pp.partial && pp.side) {
    T* side = reinterpret_cast<T*>(pp.side);
    for (int j = blockIdx.x; j < pp.side_c / 8; j += gridDim.x)
      for (int rb = threadIdx.x; rb < rows; rb += SMALL_U * SMALL_THREADS) {
        float8 pa[SMALL_U];
#pragma unroll
        for (int u = 0; u < SMALL_U; ++u) {
          const int r = rb + u * SMALL_THREADS;
          if (r < rows) pa[u] = ld8(pp.partial + (size_t)r * pp.ld + 8 * j);
        }
#pragma unroll
        for (int u = 0; u < SMALL_U; ++u) {
          const int r = rb + u * SMALL_THREADS;
          if (r < rows) {
            st8(pp.partial + (size_t)r * pp.ld + 8 * j, float8_zero());
            straw8(side + (size_t)r * pp.side_c + 8 * j, round_raw8(pa[u], (const T*)nullptr));
          }
        }
      }
  }
  for (int rb = threadIdx.x; rb < rows; rb += SMALL_U * SMALL_THREADS) {
    R8 qx[SMALL_U], qa[SMALL_U], qb[SMALL_U];
    float8 pa[SMALL_U];
#pragma unroll
    for (int u = 0; u < SMALL_U; ++u) {
      const int r = rb + u * SMALL_THREADS;
      if (r < rows) {
        const size_t off = (size_t)r * C + c0;
        qx[u] = ldraw8(x + off);
        if (pp.partial) pa[u] = ld8(pp.partial + (size_t)r * pp.ld + pp.off + c0);
        else if (gA) qa[u] = ldraw8(gA + off);
        if (gB) qb[u] = ldraw8(gB + off);
      }
    }
#pragma unroll
    for (int u = 0; u < SMALL_U; ++u) {
      const int r = rb + u * SMALL_THREADS;
      if (r < rows) {
        GzIn in;
        sx[r] = qx[u];
        in.x = cvt8(qx[u]);
        if (pp.partial) {
          qa[u] = round_raw8(pa[u], (const T*)nullptr);
          st8(pp.partial + (size_t)r * pp.ld + pp.off + c0, float8_zero());
        }
        if (hasA) { sg[r] = qa[u]; in.a = cvt8(qa[u]); }
        if (gB) { sb[r] = qb[u]; in.b = cvt8(qb[u]); }
        const float8 g = gz_compute(in, &sc, &sh, hasA, slope0, gB != nullptr, slope1);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] += g.v[i]; v[8 + i] = fmaf(g.v[i], (in.x.v[i] - mu.v[i]) * is.v[i], v[8 + i]); }
      }
    }
  }
  double tot[16];
  slab_block_reduce16<T>(v, tot);
  if (threadIdx.x < 8) {
    const int c = c0 + threadIdx.x;
    if (dgamma) { dbeta[c] = (float)tot[threadIdx.x]; dgamma[c] = (float)tot[8 + threadIdx.x]; }
    if (sums) { sums[c] = tot[threadIdx.x]; sums[C + c] = tot[8 + threadIdx.x]; }     // (kept: eval-mode centring fix reads them)
  }
  const float inv_m = 1.f / (float)rows;
  float cA[8], cB[8], cC[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cA[i] = sc.v[i]; cB[i] = 0.f; cC[i] = 0.f;
    if (mode == 2) {
      const float s1 = (float)tot[i] * inv_m, s2 = (float)tot[8 + i] * inv_m;
      cB[i] = -sc.v[i] * is.v[i] * s2;
      cC[i] = -sc.v[i] * s1 - cB[i] * mu.v[i];
    }
  }
  for (int r = threadIdx.x; r < rows; r += SMALL_THREADS) {
    GzIn in;
    in.x = cvt8(sx[r]);
    if (hasA) in.a = cvt8(sg[r]);
    if (gB) in.b = cvt8(sb[r]);
    float8 g = gz_compute(in, &sc, &sh, hasA, slope0, gB != nullptr, slope1);
#pragma unroll
    for (int i = 0; i < 8; ++i) g.v[i] = fmaf(cA[i], g.v[i], fmaf(cB[i], in.x.v[i], cC[i]));
    st8(dx + (size_t)r * C + c0, g);
  }
}

// ------------------------------------------------------------------ output head backward
__global__ void __launch_bounds__(EW_THREADS)
head_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, long long n, int final_sigmoid,
                float* __restrict__ du, float* __restrict__ dbias) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float yv = y[i], g = dy[i];
    float d = final_sigmoid ? g * yv * (1.f - yv) : (yv > 0.f ? g : 0.f);
    du[i] = d;
    acc += d;
  }
  acc = warp_sum(acc);
  __shared__ float red[EW_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < EW_THREADS / 32; ++w) t += red[w];
    atomicAdd(dbias, t);
  }
}

__global__ void __launch_bounds__(EW_THREADS)
cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    st4(dst + 4 * i, ld4(src + 4 * i));
}

// src [R][16][C] fp32 -> dst [C][16][R] bf16, 32x32 tiles per tap through shared memory
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int R, int C) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < R && c < C) ? src[((size_t)r * 16 + tap) * C + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    int c = c0 + y, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[((size_t)c * 16 + tap) * R + r] = __float2bfloat16_rn(tile[threadIdx.x][y]);
  }
}

int ew_grid(long long n4) {
  long long blocks = (n4 + EW_THREADS - 1) / EW_THREADS;
  long long cap = (long long)adp::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct ColLaunch { dim3 grid, block; };
ColLaunch col_launch8(long long rows, int C) {
  int tx = C / 8 < 32 ? C / 8 : 32;
  int p2 = 1;
  while (p2 * 2 <= tx) p2 *= 2;       // power of two so that 256 / tx is integral
  tx = p2;
  const int ty = 256 / tx;
  const int gx = adp_cdiv(C, tx * 8);
  long long gy = (rows + (long long)ty * 8 - 1) / ((long long)ty * 8);   // >= 8 rows per thread
  static const int per_sm = getenv("ADP_COL_BLOCKS_PER_SM") ? atoi(getenv("ADP_COL_BLOCKS_PER_SM")) : 2;   // fewer blocks = fewer fp64 atomics per channel cell (4 -> 2: -33 us / step)
  long long cap = (long long)adp::sm_count() * per_sm / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return ColLaunch{dim3(gx, (unsigned)gy), dim3(tx, ty)};
}

dim3 col_grid(long long rows, int C) {
  int gx = adp_cdiv(C, 128);
  long long gy = (rows + 63) / 64;  // >= 8 rows per thread
  long long cap = (long long)adp::sm_count() * 4 / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}

}  // namespace

namespace adp {

#define ADP_DISPATCH_T(dtype, ...)                          \
  if ((dtype) == ADP_F32) {                                 \
    using T = float;                                        \
    __VA_ARGS__                                             \
  } else if ((dtype) == ADP_BF16) {                         \
    using T = bf16;                                         \
    __VA_ARGS__                                             \
  } else {                                                  \
    adp_set_error("unknown dtype %d", (int)(dtype));        \
    return ADP_ERR_ARG;                                     \
  }

bool bn_small_ok(int dtype, long long rows, int C, int tensors) {
  static const int on = getenv("ADP_BN_SMALL") ? atoi(getenv("ADP_BN_SMALL")) : 1;
  // (ADP_BN_SMALL_ROWS: 1024 rows was the break-even of C/8 blocks with one row per thread in flight; with four in flight
  // and the split-K finish folded in, 4096 rows -- the 8x8 levels at B = 64 -- is 30 us per step ahead of three launches)
  static const int max_rows = getenv("ADP_BN_SMALL_ROWS") ? atoi(getenv("ADP_BN_SMALL_ROWS")) : 4096;
  const size_t per_row = dtype == ADP_F32 ? 32 : 16;
  // (above 1024 rows only with >= 64 blocks: a 4096-row slab per block is a long serial walk, it needs the width)
  if (rows > 1024 && C < 512) return false;
  return on && C % 8 == 0 && rows >= 1 && rows <= max_rows && (size_t)rows * per_row * tensors <= 200 * 1024;
}

int bn_small_fwd(int dtype, const void* x, long long rows, int C, const BnFin& f, float slope0, void* out0, float slope1,
                 void* out1, cudaStream_t s, float* partial) {
  const int smem = (int)(rows * (dtype == ADP_F32 ? 32 : 16));
  ADP_SMEM_ATTR(bn_small_fwd_kernel<float>, 200 * 1024);
  ADP_SMEM_ATTR(bn_small_fwd_kernel<bf16>, 200 * 1024);
  ADP_DISPATCH_T(dtype, bn_small_fwd_kernel<T><<<C / 8, SMALL_THREADS, smem, s>>>((const T*)x, (int)rows, C, f,
                                       slope0, (T*)out0, slope1, (T*)out1, partial, (T*)const_cast<void*>(x));)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int bn_small_bwd(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift, const float* mean,
                 const float* invstd, const void* gA, float slope0, const void* gB, float slope1, int mode, void* dx,
                 float* dgamma, float* dbeta, double* sums, cudaStream_t s, const BnSmallPartial* pp) {
  const int smem = (int)(rows * (dtype == ADP_F32 ? 32 : 16) * 3);
  BnSmallPartial part;
  memset(&part, 0, sizeof(part));
  if (pp) part = *pp;
  ADP_CHECK_ARG(!part.partial || (part.ld % 8 == 0 && part.off % 8 == 0 && part.side_c % 8 == 0 && part.off + C <= part.ld &&
                                  part.side_c <= part.off), "bn_small_bwd: bad partial-sum layout");
  ADP_SMEM_ATTR(bn_small_bwd_kernel<float>, 200 * 1024);
  ADP_SMEM_ATTR(bn_small_bwd_kernel<bf16>, 200 * 1024);
  ADP_DISPATCH_T(dtype, bn_small_bwd_kernel<T><<<C / 8, SMALL_THREADS, smem, s>>>((const T*)x, (int)rows, C,
                                       scale, shift, mean, invstd, (const T*)gA, slope0, (const T*)gB, slope1, mode, (T*)dx,
                                       dgamma, dbeta, sums, part);)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int finish_act(int dtype, float* partial, long long n, const void* mask, float slope, void* raw, void* act, cudaStream_t s) {
  ADP_CHECK_ARG(partial && act && (mask || raw) && n > 0 && n % 8 == 0, "finish_act: bad arguments");
  const long long n8 = n / 8;
  ADP_DISPATCH_T(dtype, finish_act8_kernel<T><<<ew_grid(n8), EW_THREADS, 0, s>>>(partial, n8, (const T*)mask, slope, (T*)raw, (T*)act);)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int bn_stats(int dtype, const void* x, long long rows, int C, double* sums, cudaStream_t s) {
  ADP_CHECK_ARG(C % 4 == 0, "bn_stats: C %% 4 != 0");
  if (C % 8 == 0) {
    ColLaunch L = col_launch8(rows, C);
    ADP_DISPATCH_T(dtype, bn_stats8_kernel<T><<<L.grid, L.block, 0, s>>>((const T*)x, rows, C, sums);)
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  }
  ADP_DISPATCH_T(dtype, bn_stats_kernel<T><<<col_grid(rows, C), dim3(32, 8), 0, s>>>((const T*)x, rows, C, sums);)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int bn_finalize(const BnFin& f, int C, cudaStream_t s) {
  bn_finalize_kernel<<<adp_cdiv(C, 128), 128, 0, s>>>(f, C);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int bn_affine_act(int dtype, const void* x, long long rows, int C, const BnFin& f, float slope0, void* out0, float slope1,
                  void* out1, cudaStream_t s) {
  if (C % 8 == 0 && 2048 % C == 0) {
    const long long n8 = rows * C / 8;
    ADP_DISPATCH_T(dtype, (bn_affine_act8_kernel<T><<<ew_grid(n8), EW_THREADS, 0, s>>>((const T*)x, n8, C, f, slope0, (T*)out0,
                                                                                     slope1, (T*)out1));)
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  }
  ADP_TRY(bn_finalize(f, C, s));
  return affine_act(dtype, x, rows, C, f.scale, f.shift, slope0, out0, slope1, out1, s);
}

int affine_act(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
               float slope0, void* out0, float slope1, void* out1, cudaStream_t s) {
  ADP_CHECK_ARG(C % 4 == 0, "affine_act: C %% 4 != 0");
  if (C % 8 == 0) {
    long long n8 = rows * C / 8;
    if (2048 % C == 0) {
      ADP_DISPATCH_T(dtype, (affine_act8_kernel<T, true><<<ew_grid(n8), EW_THREADS, 0, s>>>(
                                (const T*)x, n8, C, scale, shift, slope0, (T*)out0, slope1, (T*)out1));)
    } else {
      ADP_DISPATCH_T(dtype, (affine_act8_kernel<T, false><<<ew_grid(n8), EW_THREADS, 0, s>>>(
                                (const T*)x, n8, C, scale, shift, slope0, (T*)out0, slope1, (T*)out1));)
    }
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  }
  long long n4 = rows * C / 4;
  ADP_DISPATCH_T(dtype, affine_act_kernel<T><<<ew_grid(n4), EW_THREADS, 0, s>>>(
                            (const T*)x, n4, C, scale, shift, slope0, (T*)out0, slope1, (T*)out1);)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int act_bn_bwd_reduce(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const void* gA, float slope0, const void* gB,
                      float slope1, double* sums, cudaStream_t s) {
  ADP_CHECK_ARG(C % 4 == 0, "act_bn_bwd_reduce: C %% 4 != 0");
  if (C % 8 == 0) {
    ColLaunch L = col_launch8(rows, C);
    const size_t smem = (size_t)RED_STAGES * RED_ROWS * 3 * 256 * (dtype == ADP_F32 ? 32 : 16);
    ADP_SMEM_ATTR(act_bn_bwd_reduce8_kernel<float>, RED_STAGES * RED_ROWS * 3 * 256 * 32);
    ADP_SMEM_ATTR(act_bn_bwd_reduce8_kernel<bf16>, RED_STAGES * RED_ROWS * 3 * 256 * 16);
    ADP_DISPATCH_T(dtype, act_bn_bwd_reduce8_kernel<T><<<L.grid, L.block, smem, s>>>(
                              (const T*)x, rows, C, scale, shift, mean, invstd, (const T*)gA, slope0, (const T*)gB,
                              slope1, sums);)
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  }
  ADP_DISPATCH_T(dtype, act_bn_bwd_reduce_kernel<T><<<col_grid(rows, C), dim3(32, 8), 0, s>>>(
                            (const T*)x, rows, C, scale, shift, mean, invstd, (const T*)gA, slope0,
                            (const T*)gB, slope1, sums);)
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int act_bn_bwd_apply(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
                     const float* mean, const float* invstd, const void* gA, float slope0, const void* gB,
                     float slope1, const double* sums, int mode, void* dx, float* dgamma, float* dbeta, cudaStream_t s) {
  ADP_CHECK_ARG(C % 4 == 0, "act_bn_bwd_apply: C %% 4 != 0");
  ADP_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr) && (!dgamma || sums), "act_bn_bwd_apply: dgamma/dbeta need sums");
  if (C % 8 == 0) {
    long long n8 = rows * C / 8;
    if (2048 % C == 0) {
      const size_t smem = (size_t)RED_STAGES * RED_ROWS * 3 * 256 * (dtype == ADP_F32 ? 32 : 16);
      ADP_SMEM_ATTR((act_bn_bwd_apply8_kernel<float, true>), RED_STAGES * RED_ROWS * 3 * 256 * 32);
      ADP_SMEM_ATTR((act_bn_bwd_apply8_kernel<bf16, true>), RED_STAGES * RED_ROWS * 3 * 256 * 16);
      ADP_DISPATCH_T(dtype, (act_bn_bwd_apply8_kernel<T, true><<<ew_grid(n8), EW_THREADS, smem, s>>>(
                                (const T*)x, n8, rows, C, scale, shift, mean, invstd, (const T*)gA, slope0,
                                (const T*)gB, slope1, sums, mode, (T*)dx, dgamma, dbeta));)
    } else {
      ADP_DISPATCH_T(dtype, (act_bn_bwd_apply8_kernel<T, false><<<ew_grid(n8), EW_THREADS, 0, s>>>(
                                (const T*)x, n8, rows, C, scale, shift, mean, invstd, (const T*)gA, slope0,
                                (const T*)gB, slope1, sums, mode, (T*)dx, dgamma, dbeta));)
    }
    ADP_LAUNCH_CHECK();
    return ADP_OK;
  }
  long long n4 = rows * C / 4;
  ADP_DISPATCH_T(dtype, act_bn_bwd_apply_kernel<T><<<ew_grid(n4), EW_THREADS, 0, s>>>(
                            (const T*)x, n4, rows, C, scale, shift, mean, invstd, (const T*)gA, slope0,
                            (const T*)gB, slope1, sums, mode, (T*)dx);)
  ADP_LAUNCH_CHECK();
  if (dgamma) return bn_param_grads(sums, C, dgamma, dbeta, s);
  return ADP_OK;
}

int bn_param_grads(const double* sums, int C, float* dgamma, float* dbeta, cudaStream_t s) {
  bn_param_grads_kernel<<<adp_cdiv(C, 128), 128, 0, s>>>(sums, C, dgamma, dbeta);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int head_bwd(const float* y, const float* dy, long long n, int final_sigmoid, float* du, float* dbias,
             cudaStream_t s) {
  long long blocks = (n + EW_THREADS * 8 - 1) / (EW_THREADS * 8);
  long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  head_bwd_kernel<<<(int)blocks, EW_THREADS, 0, s>>>(y, dy, n, final_sigmoid, du, dbias);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
  ADP_CHECK_ARG(n % 4 == 0, "cast: n %% 4 != 0");
  cast_kernel<<<ew_grid(n / 4), EW_THREADS, 0, s>>>(src, (bf16*)dst, n / 4);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int cast_transpose_taps(const float* src, void* dst, int R, int C, cudaStream_t s) {
  dim3 grid(adp_cdiv(C, 32), adp_cdiv(R, 32), 16);
  cast_transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(src, (bf16*)dst, R, C);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

}  // namespace adp
