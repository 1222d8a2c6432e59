// Building blocks of the binaural-attention network (BASELINE config 4):
//   models/binaural_attention_model.py:22-39  DoubleConv   (3x3 s1 p1 conv, BatchNorm, ReLU) x 2
//   :42-53 Down (MaxPool2d(2)), :56-78 Up (bilinear x2, align_corners=True, concat), :81-153 cross attention,
//   :240-247 fusion 1x1 convs, :262-265 output head.
// Activations are bf16 NHWC; the dense contractions run on the tcgen05 implicit-GEMM kernels of adp_conv_tc.cu /
// adp_wgrad_tc.cu (mode 4 = 3x3, mode 2 = row GEMM), everything here is the bandwidth-bound glue.
#include <math.h>
#include "adp_common.cuh"

using namespace adp;

extern "C" int adp_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  ADP_CHECK_ARG(src && dst && n > 0, "cast_bf16: bad arguments");
  return cast_f32_to_bf16(src, dst, (long long)n, (cudaStream_t)stream);
}

extern "C" int adp_conv2d_k3s1_fprop(const void* x0, int C0, const void* x1, int C1, const void* w_bf16, void* y, int B, int H,
                                     int W, int Cout, void* scratch, size_t scratch_bytes, void* stream) {
  ADP_CHECK_ARG(x0 && w_bf16 && y && (C1 == 0 || x1), "conv2d_k3s1_fprop: null pointer");
  ADP_CHECK_ARG(tc_supported_conv3x3(B, H, W, C0, C1, Cout, 0),
                "conv2d_k3s1_fprop: unsupported shape B=%d %dx%d C=%d+%d -> %d (needs sm_100, power-of-two H, W, channels %% 64)",
                B, H, W, C0, C1, Cout);
  return tc_conv3x3(x0, C0, x1, C1, w_bf16, 0, y, Cout, nullptr, 0, B, H, W, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int adp_conv2d_k3s1_dgrad(const void* dy, int Cout, const void* w_bf16, void* dx0, int C0, void* dx1, int C1, int B,
                                     int H, int W, void* scratch, size_t scratch_bytes, void* stream) {
  ADP_CHECK_ARG(dy && w_bf16 && dx0 && (C1 == 0 || dx1), "conv2d_k3s1_dgrad: null pointer");
  ADP_CHECK_ARG(tc_supported_conv3x3(B, H, W, Cout, 0, C0, C1), "conv2d_k3s1_dgrad: unsupported shape B=%d %dx%d %d -> %d+%d", B,
                H, W, Cout, C0, C1);
  return tc_conv3x3(dy, Cout, nullptr, 0, w_bf16, 1, dx0, C0, dx1, C1, B, H, W, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int adp_conv2d_k3s1_wgrad(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw, int B,
                                     int H, int W, void* stream) {
  ADP_CHECK_ARG(dy && x0 && dw && (C1 == 0 || x1), "conv2d_k3s1_wgrad: null pointer");
  ADP_CHECK_ARG(tc_supported_wgrad3x3(B, H, W, Cout, C0) && (C1 == 0 || tc_supported_wgrad3x3(B, H, W, Cout, C1)),
                "conv2d_k3s1_wgrad: unsupported shape B=%d %dx%d %d x (%d+%d)", B, H, W, Cout, C0, C1);
  cudaStream_t s = (cudaStream_t)stream;
  ADP_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 9 * (size_t)Cout * (C0 + C1), s));
  ADP_TRY(tc_wgrad3x3(dy, Cout, x0, C0, C0 + C1, 0, dw, B, H, W, s));
  if (C1 > 0) ADP_TRY(tc_wgrad3x3(dy, Cout, x1, C1, C0 + C1, C0, dw, B, H, W, s));
  return ADP_OK;
}

extern "C" int adp_gemm_rows_bf16(const void* a0, int K0, const void* a1, int K1, const void* b, int b_kn, void* c_bf16_0,
                                  int N0, void* c_bf16_1, int N1, float* c_f32, int64_t M, void* stream) {
  ADP_CHECK_ARG(a0 && b && (K1 == 0 || a1), "gemm_rows_bf16: null pointer");
  return tc_gemm_rows(a0, K0, a1, K1, b, b_kn, c_bf16_0, N0, c_bf16_1, N1, c_f32, (long long)M, (cudaStream_t)stream);
}

// =====================================================================================================================
// bandwidth-bound glue kernels (bf16 NHWC rows, 8 channels = 16 bytes per thread)
namespace {

constexpr int BT = 256;
inline int grid_for(long long items, int per_block = BT) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- first conv of an encoder: one input channel (a plane of the [B,2,H,W] fp32 network input) -> Cout ------------
__global__ void __launch_bounds__(BT)
conv3x3_c1_fprop_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ w, bf16* __restrict__ y, int B,
                        int H, int W, int Cout) {
  extern __shared__ float ws[];                       // [9][Cout]  (transposed so that 8 channels are contiguous)
  for (int i = threadIdx.x; i < 9 * Cout; i += BT) ws[(i % 9) * Cout + i / 9] = w[i];
  __syncthreads();
  const int g8 = Cout / 8;
  const long long items = (long long)B * H * W * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    const long long pix = it / g8;
    const int xx = (int)(pix % W), yy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const float* xp = x + (long long)b * xbs;
    float8 acc;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
      const float v = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? xp[(long long)sy * W + sx] : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc.v[k] = fmaf(v, ws[t * Cout + c0 + k], acc.v[k]);
    }
    st8(y + pix * Cout + c0, acc);
  }
}

// dw[co][tap] += sum_pix dy[pix][co] * x[pix + shift(tap)]
__global__ void __launch_bounds__(BT)
conv3x3_c1_wgrad_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, long long xbs, float* __restrict__ dw, int B,
                        int H, int W, int Cout) {
  extern __shared__ float acc_s[];                    // [Cout][9]
  for (int i = threadIdx.x; i < 9 * Cout; i += BT) acc_s[i] = 0.f;
  __syncthreads();
  const int g8 = Cout / 8;
  const int cg = threadIdx.x % g8, lane_p = threadIdx.x / g8, pl = BT / g8;
  float acc[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
  const long long npix = (long long)B * H * W;
  if (lane_p < pl) {
    for (long long pix = (long long)blockIdx.x * pl + lane_p; pix < npix; pix += (long long)gridDim.x * pl) {
      const int xx = (int)(pix % W), yy = (int)((pix / W) % H);
      const int b = (int)(pix / ((long long)W * H));
      const float* xp = x + (long long)b * xbs;
      const float8 g = ld8(dy + pix * Cout + cg * 8);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int sy = yy + t / 3 - 1, sx = xx + t % 3 - 1;
        const float v = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? xp[(long long)sy * W + sx] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k][t] = fmaf(g.v[k], v, acc[k][t]);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&acc_s[(cg * 8 + k) * 9 + t], acc[k][t]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * Cout; i += BT) atomicAdd(&dw[i], acc_s[i]);
}

// ---- MaxPool2d(2) ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
maxpool2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int Ho, int Wo, int C) {
  const int g8 = C / 8;
  const long long items = (long long)B * Ho * Wo * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    const long long opix = it / g8;
    const int j = (int)(opix % Wo), i = (int)((opix / Wo) % Ho);
    const long long b = opix / ((long long)Wo * Ho);
    const bf16* p00 = x + ((b * 2 * Ho + 2 * i) * (2LL * Wo) + 2 * j) * C + c0;
    const float8 a = ld8(p00), bq = ld8(p00 + C), c = ld8(p00 + 2LL * Wo * C), d = ld8(p00 + 2LL * Wo * C + C);
    float8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaxf(a.v[k], bq.v[k]), fmaxf(c.v[k], d.v[k]));
    st8(y + opix * C + c0, o);
  }
}

// the gradient goes to the first maximum in scan order (torch: `val > maxval`), recomputed from the saved input
__global__ void __launch_bounds__(BT)
maxpool2_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int Ho, int Wo, int C) {
  const int g8 = C / 8;
  const long long items = (long long)B * Ho * Wo * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    const long long opix = it / g8;
    const int j = (int)(opix % Wo), i = (int)((opix / Wo) % Ho);
    const long long b = opix / ((long long)Wo * Ho);
    const long long base = ((b * 2 * Ho + 2 * i) * (2LL * Wo) + 2 * j) * C + c0;
    const long long off[4] = {0, C, 2LL * Wo * C, 2LL * Wo * C + C};
    float8 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = ld8(x + base + off[q]);
    const float8 g = ld8(dy + opix * C + c0);
    float8 o[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int arg = 0;
      float m = v[0].v[k];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q].v[k] > m) { m = v[q].v[k]; arg = q; }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q].v[k] = q == arg ? g.v[k] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st8(dx + base + off[q], o[q]);
  }
}

// ---- nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) -------------------------------------------------
struct Tap2 { int i0, i1; float w0, w1; };
__device__ __forceinline__ Tap2 up_tap(int dst, int n_in, int n_out) {
  // aten area_pixel_compute_source_index with align_corners: src = dst * (in - 1) / (out - 1)
  const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
  const float src = scale * (float)dst;
  Tap2 t;
  t.i0 = (int)src;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.w1 = src - (float)t.i0;
  t.w0 = 1.f - t.w1;
  return t;
}

__global__ void __launch_bounds__(BT)
upsample2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  const int g8 = C / 8, Ho = 2 * H, Wo = 2 * W;
  const long long items = (long long)B * Ho * Wo * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    const long long opix = it / g8;
    const int j = (int)(opix % Wo), i = (int)((opix / Wo) % Ho);
    const long long b = opix / ((long long)Wo * Ho);
    const Tap2 ty = up_tap(i, H, Ho), tx = up_tap(j, W, Wo);
    const bf16* xb = x + b * H * W * C + c0;
    const float8 a = ld8(xb + ((long long)ty.i0 * W + tx.i0) * C), bq = ld8(xb + ((long long)ty.i0 * W + tx.i1) * C),
                 c = ld8(xb + ((long long)ty.i1 * W + tx.i0) * C), d = ld8(xb + ((long long)ty.i1 * W + tx.i1) * C);
    float8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o.v[k] = ty.w0 * (tx.w0 * a.v[k] + tx.w1 * bq.v[k]) + ty.w1 * (tx.w0 * c.v[k] + tx.w1 * d.v[k]);
    st8(y + opix * C + c0, o);
  }
}

// exact adjoint in gather form: every source pixel collects the (<= 5 x 5) destination pixels that read it
__global__ void __launch_bounds__(BT)
upsample2_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int H, int W, int C) {
  const int g8 = C / 8, Ho = 2 * H, Wo = 2 * W;
  const long long items = (long long)B * H * W * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    const long long pix = it / g8;
    const int sx = (int)(pix % W), sy = (int)((pix / W) % H);
    const long long b = pix / ((long long)W * H);
    float8 acc;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
    const int ylo = max(2 * sy - 3, 0), yhi = min(2 * sy + 3, Ho - 1);
    const int xlo = max(2 * sx - 3, 0), xhi = min(2 * sx + 3, Wo - 1);
    for (int i = ylo; i <= yhi; ++i) {
      const Tap2 ty = up_tap(i, H, Ho);
      const float wy = (ty.i0 == sy ? ty.w0 : 0.f) + (ty.i1 == sy ? ty.w1 : 0.f);
      if (wy == 0.f) continue;
      for (int j = xlo; j <= xhi; ++j) {
        const Tap2 tx = up_tap(j, W, Wo);
        const float wx = (tx.i0 == sx ? tx.w0 : 0.f) + (tx.i1 == sx ? tx.w1 : 0.f);
        if (wx == 0.f) continue;
        const float8 g = ld8(dy + ((b * Ho + i) * Wo + j) * C + c0);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.v[k] = fmaf(wy * wx, g.v[k], acc.v[k]);
      }
    }
    st8(dx + pix * C + c0, acc);
  }
}

// ---- rows [R][C]: bias, residual, column sums, row dots -----------------------------------------------------------------
__global__ void __launch_bounds__(BT)
bias_rows_kernel(bf16* __restrict__ x, const float* __restrict__ bias, long long rows, int C) {
  const int g8 = C / 8;
  const long long items = rows * g8;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < items; it += (long long)gridDim.x * BT) {
    const int c0 = (int)(it % g8) * 8;
    float8 v = ld8(x + it * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] += bias[c0 + k];
    st8(x + it * 8, v);
  }
}

// y = a + gamma[0] * b   (left_out = left_feat + gamma * attended, :134)
__global__ void __launch_bounds__(BT)
axpy_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const float* __restrict__ gamma, bf16* __restrict__ y,
                 long long n8) {
  const float g = gamma[0];
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n8; it += (long long)gridDim.x * BT) {
    const float8 u = ld8(a + it * 8), v = ld8(b + it * 8);
    float8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaf(g, v.v[k], u.v[k]);
    st8(y + it * 8, o);
  }
}

// out[0] += sum a*b (fp32 atomics on block partials): dgamma = sum(dy * attended);  scaled copy: y = s[0] * a
__global__ void __launch_bounds__(BT)
dot_all_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long n8, float* __restrict__ out) {
  float acc = 0.f;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n8; it += (long long)gridDim.x * BT) {
    const float8 u = ld8(a + it * 8), v = ld8(b + it * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(u.v[k], v.v[k], acc);
  }
  acc = warp_sum(acc);
  __shared__ float red[BT / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < BT / 32; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(BT)
scale_rows_kernel(const bf16* __restrict__ a, const float* __restrict__ s, bf16* __restrict__ y, long long n8) {
  const float g = s[0];
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n8; it += (long long)gridDim.x * BT) {
    float8 u = ld8(a + it * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) u.v[k] *= g;
    st8(y + it * 8, u);
  }
}

// y = a + b (gradient fan-in of a tensor with two consumers)
__global__ void __launch_bounds__(BT)
add_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, long long n8) {
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n8; it += (long long)gridDim.x * BT) {
    const float8 u = ld8(a + it * 8), v = ld8(b + it * 8);
    float8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = u.v[k] + v.v[k];
    st8(y + it * 8, o);
  }
}

__global__ void sums_to_float_kernel(const double* __restrict__ sums, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)sums[c];
}

// out[r] = sum_c a[r][c] * b[r][c]   (one warp per row)
__global__ void __launch_bounds__(BT)
rowdot_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long rows, int C, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (BT / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (BT / 32)) {
    float acc = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      const float8 u = ld8(a + r * C + c), v = ld8(b + r * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(u.v[k], v.v[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
  }
}

// ---- attention softmax pieces (scores fp32 [R][N], row = query) ---------------------------------------------------------
// P[r][:] = softmax(scale * S[r][:]) in bf16; m[r] = max(scale*S), l[r] = sum exp(scale*S - m)
__global__ void __launch_bounds__(BT)
softmax_rows_kernel(const float* __restrict__ S, long long R, int N, float scale, bf16* __restrict__ P, float* __restrict__ m_out,
                    float* __restrict__ l_out) {
  __shared__ float red[BT / 32];
  __shared__ float bcast;
  for (long long r = blockIdx.x; r < R; r += gridDim.x) {
    const float* row = S + r * (long long)N;
    float mx = -INFINITY;
    for (int c = threadIdx.x * 4; c < N; c += BT * 4) {
      const float4 v = ld4(row + c);
      mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = red[0];
      for (int w = 1; w < BT / 32; ++w) t = fmaxf(t, red[w]);
      bcast = t * scale;
    }
    __syncthreads();
    const float m = bcast;
    float sum = 0.f;
    for (int c = threadIdx.x * 4; c < N; c += BT * 4) {
      const float4 v = ld4(row + c);
      sum += __expf(v.x * scale - m) + __expf(v.y * scale - m) + __expf(v.z * scale - m) + __expf(v.w * scale - m);
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < BT / 32; ++w) t += red[w];
      bcast = t;
      m_out[r] = m;
      l_out[r] = t;
    }
    __syncthreads();
    const float inv = 1.f / bcast;
    for (int c = threadIdx.x * 4; c < N; c += BT * 4) {
      const float4 v = ld4(row + c);
      st4(P + r * (long long)N + c, make_float4(__expf(v.x * scale - m) * inv, __expf(v.y * scale - m) * inv,
                                                __expf(v.z * scale - m) * inv, __expf(v.w * scale - m) * inv));
    }
    __syncthreads();
  }
}

// by_col = 0: P[r][c] = exp(scale*S[r][c] - m[r]) / l[r];  by_col = 1 (transposed scores): statistics indexed by c
__global__ void __launch_bounds__(BT)
softmax_apply_kernel(const float* __restrict__ S, long long R, int N, float scale, const float* __restrict__ m,
                     const float* __restrict__ l, int by_col, bf16* __restrict__ P) {
  const long long n4 = R * (long long)N / 4;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n4; it += (long long)gridDim.x * BT) {
    const long long e = it * 4;
    const long long r = e / N;
    const int c = (int)(e - r * N);
    const float4 v = ld4(S + e);
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long si = by_col ? c + k : r;
      o[k] = __expf(o[k] * scale - m[si]) / l[si];
    }
    st4(P + e, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// dS = scale * P * (dP - delta[query]);  by_col selects whether the query index is the column (transposed form)
__global__ void __launch_bounds__(BT)
softmax_bwd_kernel(const bf16* __restrict__ P, const float* __restrict__ dP, long long R, int N, float scale,
                   const float* __restrict__ delta, int by_col, bf16* __restrict__ dS) {
  const long long n4 = R * (long long)N / 4;
  for (long long it = (long long)blockIdx.x * BT + threadIdx.x; it < n4; it += (long long)gridDim.x * BT) {
    const long long e = it * 4;
    const long long r = e / N;
    const int c = (int)(e - r * N);
    const float4 g = ld4(dP + e);
    const uint2 pu = *reinterpret_cast<const uint2*>(P + e);
    const float p[4] = {__uint_as_float(pu.x << 16), __uint_as_float(pu.x & 0xffff0000u), __uint_as_float(pu.y << 16),
                        __uint_as_float(pu.y & 0xffff0000u)};
    const float gg[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = scale * p[k] * (gg[k] - delta[by_col ? c + k : r]);
    st4(dS + e, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ---- output head: Conv2d(C, 1, 1) + Sigmoid, * max_depth, clamp (:262-265, :318-332) ------------------------------------
__global__ void __launch_bounds__(BT)
head_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float max_depth,
                long long rows, int C, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (BT / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (BT / 32)) {
    float acc = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
      const float8 u = ld8(x + r * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(u.v[k], w[c + k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float sg = 1.f / (1.f + __expf(-(acc + bias[0])));
      y[r] = fminf(fmaxf(sg * max_depth, 0.f), max_depth);
    }
  }
}

// du = dy * max_depth * s(1-s);  dx[r][c] = du * w[c];  dw[c] += sum_r du * x[r][c];  db += sum du
__global__ void __launch_bounds__(BT)
head_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float max_depth,
                const float* __restrict__ dy, long long rows, int C, bf16* __restrict__ dx, float* __restrict__ dw,
                float* __restrict__ db) {
  extern __shared__ float dw_s[];                     // [C] + 1
  for (int i = threadIdx.x; i <= C; i += BT) dw_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float dwl[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) dwl[k] = 0.f;
  float dbl = 0.f;
  for (long long r = (long long)blockIdx.x * (BT / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (BT / 32)) {
    float acc = 0.f;
    float8 u;
    const int c = lane * 8;                           // C <= 256: one 8-channel group per lane
    if (c < C) {
      u = ld8(x + r * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(u.v[k], w[c + k], acc);
    }
    acc = warp_sum(acc);
    const float sg = 1.f / (1.f + __expf(-(acc + bias[0])));
    const float du = dy[r] * max_depth * sg * (1.f - sg);
    if (c < C) {
      float8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) { o.v[k] = du * w[c + k]; dwl[k] = fmaf(du, u.v[k], dwl[k]); }
      st8(dx + r * C + c, o);
    }
    if (lane == 0) dbl += du;
  }
  if (lane * 8 < C) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&dw_s[lane * 8 + k], dwl[k]);
  }
  if (lane == 0) atomicAdd(&dw_s[C], dbl);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += BT) atomicAdd(&dw[i], dw_s[i]);
  if (threadIdx.x == 0) atomicAdd(db, dw_s[C]);
}

}  // namespace

namespace {
__global__ void rm_bias_kernel(float* __restrict__ rm, const float* __restrict__ bias, float k, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = rm[c] + k * bias[c];
}

// ---- F.interpolate(mode='bilinear', align_corners=False) of fp32 planes (binaural_attention_model.py:322-328) ----------------
// ATen's area_pixel_compute_source_index: src = max(0, (dst + 0.5) * in/out - 0.5); i0 = floor(src), i1 = min(i0 + 1, in - 1)
__device__ __forceinline__ void bilinear_src(int d, float scale, int n_in, int& i0, int& i1, float& l1) {
  float src = ((float)d + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
__global__ void __launch_bounds__(BT)
bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int Hi, int Wi, int Ho, int Wo,
                    float sh, float sw) {
  const long long total = planes * Ho * Wo;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(o % Wo);
    const long long r = o / Wo;
    const int oy = (int)(r % Ho);
    const long long pl = r / Ho;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(oy, sh, Hi, y0, y1, ly);
    bilinear_src(ox, sw, Wi, x0, x1, lx);
    const float* src = x + pl * Hi * Wi;
    const float top = (1.f - lx) * src[(size_t)y0 * Wi + x0] + lx * src[(size_t)y0 * Wi + x1];
    const float bot = (1.f - lx) * src[(size_t)y1 * Wi + x0] + lx * src[(size_t)y1 * Wi + x1];
    y[o] = (1.f - ly) * top + ly * bot;
  }
}
// dx zeroed by the caller
__global__ void __launch_bounds__(BT)
bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long planes, int Hi, int Wi, int Ho, int Wo,
                    float sh, float sw) {
  const long long total = planes * Ho * Wo;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(o % Wo);
    const long long r = o / Wo;
    const int oy = (int)(r % Ho);
    const long long pl = r / Ho;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(oy, sh, Hi, y0, y1, ly);
    bilinear_src(ox, sw, Wi, x0, x1, lx);
    float* dst = dx + pl * Hi * Wi;
    const float g = dy[o];
    atomicAdd(dst + (size_t)y0 * Wi + x0, (1.f - ly) * (1.f - lx) * g);
    atomicAdd(dst + (size_t)y0 * Wi + x1, (1.f - ly) * lx * g);
    atomicAdd(dst + (size_t)y1 * Wi + x0, ly * (1.f - lx) * g);
    atomicAdd(dst + (size_t)y1 * Wi + x1, ly * lx * g);
  }
}

// ---- nn.ConvTranspose2d(k2, s2) (:65-66) = one GEMM with N' = 4 N columns (a, b, n) per input pixel + this shuffle -----------
// forward: ys [B,H,W,2,2,N] (+ bias[n]) -> y [B,2H,2W,N];  backward: dy [B,2H,2W,N] -> dys [B,H,W,2,2,N]
__global__ void __launch_bounds__(BT)
pixel_shuffle2_kernel(const bf16* __restrict__ src, const float* __restrict__ bias, bf16* __restrict__ dst, int B, int H, int W,
                      int N, int inverse) {
  const int n8 = N / 8;
  const long long total = (long long)B * H * W * 4 * n8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % n8);
    long long r = i / n8;
    const int ab = (int)(r & 3);
    r >>= 2;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const long long b = r / H;
    const size_t packed = (size_t)i * 8;                                                                     // [b,y,x,a,bb,n]
    const size_t spread = ((((size_t)b * 2 * H + 2 * y + (ab >> 1)) * (2 * W)) + 2 * x + (ab & 1)) * N + c8 * 8;   // [b,2y+a,2x+bb,n]
    if (inverse) {
      *reinterpret_cast<uint4*>(dst + packed) = *reinterpret_cast<const uint4*>(src + spread);
    } else {
      float8 v = ld8(src + packed);
      if (bias) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] += bias[c8 * 8 + k];
      }
      st8(dst + spread, v);
    }
  }
}
}  // namespace

#define ADP_LAUNCH(kernel, grid, smem, s, ...)          \
  do {                                                  \
    kernel<<<(grid), BT, (smem), (s)>>>(__VA_ARGS__);   \
    ADP_LAUNCH_CHECK();                                 \
  } while (0)

extern "C" int adp_conv2d_k3s1_c1_fprop(const float* x, int64_t x_batch_stride, const float* w, void* y, int B, int H, int W,
                                        int Cout, void* stream) {
  ADP_CHECK_ARG(x && w && y && B > 0 && H > 0 && W > 0 && Cout > 0 && Cout % 8 == 0 && Cout <= 1024, "conv2d_k3s1_c1_fprop: bad arguments");
  ADP_LAUNCH(conv3x3_c1_fprop_kernel, grid_for((long long)B * H * W * (Cout / 8)), 9 * Cout * sizeof(float), (cudaStream_t)stream,
             x, (long long)x_batch_stride, w, (bf16*)y, B, H, W, Cout);
  return ADP_OK;
}

extern "C" int adp_conv2d_k3s1_c1_wgrad(const void* dy, const float* x, int64_t x_batch_stride, float* dw, int B, int H, int W,
                                        int Cout, void* stream) {
  ADP_CHECK_ARG(dy && x && dw && B > 0 && Cout % 8 == 0 && Cout / 8 <= BT && BT % (Cout / 8) == 0, "conv2d_k3s1_c1_wgrad: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  ADP_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 9 * Cout, s));
  const int pl = BT / (Cout / 8);
  long long blocks = ((long long)B * H * W + pl * 16 - 1) / (pl * 16);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  ADP_LAUNCH(conv3x3_c1_wgrad_kernel, (int)(blocks < 1 ? 1 : blocks), 9 * Cout * sizeof(float), s, (const bf16*)dy, x,
             (long long)x_batch_stride, dw, B, H, W, Cout);
  return ADP_OK;
}

// BatchNorm2d + (Leaky)ReLU over bf16 rows [rows][C] (:30-31, :33-34, :242-243).  conv_bias (optional) is the bias of a
// preceding convolution that was NOT added to x: it cancels in the normalised output and only shifts the running mean.
// saved: float [4C] = scale | shift | mean | invstd (for the backward pass); sums_ws: double [2C] scratch.
extern "C" int adp_bn_act_forward(const void* x, int64_t rows, int C, const float* gamma, const float* beta, const float* conv_bias,
                                  float* running_mean, float* running_var, int training, float eps, float momentum, float slope,
                                  void* y, float* saved, double* sums_ws, void* stream) {
  ADP_CHECK_ARG(x && y && gamma && beta && saved && sums_ws && rows > 0 && C > 0 && C % 8 == 0, "bn_act_forward: bad arguments");
  ADP_CHECK_ARG(training || (running_mean && running_var), "bn_act_forward: eval mode needs running statistics");
  cudaStream_t s = (cudaStream_t)stream;
  float* rm = running_mean;
  if (training) {
    ADP_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * C, s));
    ADP_TRY(bn_stats(ADP_BF16, x, rows, C, sums_ws, s));
  } else if (conv_bias) {      // y = ((x + b) - rm) * scale + beta: fold b into an effective running mean
    rm = reinterpret_cast<float*>(sums_ws);
    rm_bias_kernel<<<adp_cdiv(C, 128), 128, 0, s>>>(running_mean, conv_bias, -1.f, C, rm);
    ADP_LAUNCH_CHECK();
  }
  const BnFin fin{sums_ws, 1.0 / (double)rows, rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f, gamma, beta, rm,
                  running_var, training, eps, momentum, saved, saved + C, saved + 2 * C, saved + 3 * C};
  ADP_TRY(bn_affine_act(ADP_BF16, x, rows, C, fin, slope, y, 0.f, nullptr, s));
  if (training && conv_bias && running_mean) {
    rm_bias_kernel<<<adp_cdiv(C, 128), 128, 0, s>>>(running_mean, conv_bias, momentum, C, running_mean);
    ADP_LAUNCH_CHECK();
  }
  return ADP_OK;
}

extern "C" int adp_bn_act_backward(const void* x, int64_t rows, int C, const float* saved, const void* dy, float slope,
                                   int training, void* dx, float* dgamma, float* dbeta, double* sums_ws, void* stream) {
  ADP_CHECK_ARG(x && saved && dy && dx && dgamma && dbeta && sums_ws && rows > 0 && C % 8 == 0, "bn_act_backward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  ADP_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * C, s));
  ADP_TRY(act_bn_bwd_reduce(ADP_BF16, x, rows, C, saved, saved + C, saved + 2 * C, saved + 3 * C, dy, slope, nullptr, 0.f, sums_ws, s));
  return act_bn_bwd_apply(ADP_BF16, x, rows, C, saved, saved + C, saved + 2 * C, saved + 3 * C, dy, slope, nullptr, 0.f, sums_ws,
                          training ? 2 : 1, dx, dgamma, dbeta, s);
}

extern "C" int adp_maxpool2_forward(const void* x, void* y, int B, int Ho, int Wo, int C, void* stream) {
  ADP_CHECK_ARG(x && y && B > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "maxpool2_forward: bad arguments");
  ADP_LAUNCH(maxpool2_fwd_kernel, grid_for((long long)B * Ho * Wo * (C / 8)), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, B, Ho, Wo, C);
  return ADP_OK;
}
extern "C" int adp_maxpool2_backward(const void* x, const void* dy, void* dx, int B, int Ho, int Wo, int C, void* stream) {
  ADP_CHECK_ARG(x && dy && dx && B > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "maxpool2_backward: bad arguments");
  ADP_LAUNCH(maxpool2_bwd_kernel, grid_for((long long)B * Ho * Wo * (C / 8)), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)dy,
             (bf16*)dx, B, Ho, Wo, C);
  return ADP_OK;
}
extern "C" int adp_upsample2x_forward(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  ADP_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample2x_forward: bad arguments");
  ADP_LAUNCH(upsample2_fwd_kernel, grid_for((long long)B * 4 * H * W * (C / 8)), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, B, H, W, C);
  return ADP_OK;
}
extern "C" int adp_upsample2x_backward(const void* dy, void* dx, int B, int H, int W, int C, void* stream) {
  ADP_CHECK_ARG(dy && dx && B > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample2x_backward: bad arguments");
  ADP_LAUNCH(upsample2_bwd_kernel, grid_for((long long)B * H * W * (C / 8)), 0, (cudaStream_t)stream, (const bf16*)dy, (bf16*)dx, B, H, W, C);
  return ADP_OK;
}

extern "C" int adp_bilinear_resize_forward(const float* x, float* y, int64_t planes, int Hi, int Wi, int Ho, int Wo, void* stream) {
  ADP_CHECK_ARG(x && y && planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_resize_forward: bad arguments");
  ADP_LAUNCH(bilinear_fwd_kernel, grid_for((long long)planes * Ho * Wo), 0, (cudaStream_t)stream, x, y, (long long)planes, Hi, Wi,
             Ho, Wo, (float)Hi / (float)Ho, (float)Wi / (float)Wo);
  return ADP_OK;
}
extern "C" int adp_bilinear_resize_backward(const float* dy, float* dx, int64_t planes, int Hi, int Wi, int Ho, int Wo, void* stream) {
  ADP_CHECK_ARG(dy && dx && planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_resize_backward: bad arguments");
  ADP_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)planes * Hi * Wi, (cudaStream_t)stream));
  ADP_LAUNCH(bilinear_bwd_kernel, grid_for((long long)planes * Ho * Wo), 0, (cudaStream_t)stream, dy, dx, (long long)planes, Hi, Wi,
             Ho, Wo, (float)Hi / (float)Ho, (float)Wi / (float)Wo);
  return ADP_OK;
}
extern "C" int adp_pixel_shuffle2(const void* src, const float* bias, void* dst, int B, int H, int W, int N, int inverse,
                                  void* stream) {
  ADP_CHECK_ARG(src && dst && B > 0 && H > 0 && W > 0 && N > 0 && N % 8 == 0, "pixel_shuffle2: bad arguments");
  ADP_LAUNCH(pixel_shuffle2_kernel, grid_for((long long)B * H * W * 4 * (N / 8)), 0, (cudaStream_t)stream, (const bf16*)src, bias,
             (bf16*)dst, B, H, W, N, inverse);
  return ADP_OK;
}

// rows helpers.  op 0: x += bias[c] (in place, a = x);  1: y = a + g[0]*b;  2: y = a + b;  3: y = g[0]*a
extern "C" int adp_rows_op(int op, const void* a, const void* b, const float* g, void* y, int64_t rows, int C, void* stream) {
  ADP_CHECK_ARG(a && y && rows > 0 && C > 0 && C % 8 == 0 && op >= 0 && op <= 3, "rows_op: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const long long n8 = (long long)rows * C / 8;
  if (op == 0) {
    ADP_CHECK_ARG(g && a == y, "rows_op 0: needs bias and in-place operands");
    ADP_LAUNCH(bias_rows_kernel, grid_for(n8), 0, s, (bf16*)y, g, (long long)rows, C);
  } else if (op == 1) {
    ADP_CHECK_ARG(b && g, "rows_op 1: needs b and g");
    ADP_LAUNCH(axpy_rows_kernel, grid_for(n8), 0, s, (const bf16*)a, (const bf16*)b, g, (bf16*)y, n8);
  } else if (op == 2) {
    ADP_CHECK_ARG(b, "rows_op 2: needs b");
    ADP_LAUNCH(add_rows_kernel, grid_for(n8), 0, s, (const bf16*)a, (const bf16*)b, (bf16*)y, n8);
  } else {
    ADP_CHECK_ARG(g, "rows_op 3: needs g");
    ADP_LAUNCH(scale_rows_kernel, grid_for(n8), 0, s, (const bf16*)a, g, (bf16*)y, n8);
  }
  return ADP_OK;
}

// reductions.  op 0: out[c] = sum_r a[r][c] (sums_ws: double [2C]);  1: out[r] = sum_c a[r][c]*b[r][c];
//              2: out[0] = sum a*b over everything
extern "C" int adp_rows_reduce(int op, const void* a, const void* b, int64_t rows, int C, float* out, double* sums_ws, void* stream) {
  ADP_CHECK_ARG(a && out && rows > 0 && C > 0 && C % 8 == 0 && op >= 0 && op <= 2, "rows_reduce: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (op == 0) {
    ADP_CHECK_ARG(sums_ws, "rows_reduce 0: needs the double scratch");
    ADP_CUDA(cudaMemsetAsync(sums_ws, 0, sizeof(double) * 2 * C, s));
    ADP_TRY(bn_stats(ADP_BF16, a, rows, C, sums_ws, s));
    sums_to_float_kernel<<<adp_cdiv(C, 128), 128, 0, s>>>(sums_ws, C, out);
    ADP_LAUNCH_CHECK();
  } else if (op == 1) {
    ADP_CHECK_ARG(b, "rows_reduce 1: needs b");
    ADP_LAUNCH(rowdot_kernel, grid_for((long long)rows, BT / 32), 0, s, (const bf16*)a, (const bf16*)b, (long long)rows, C, out);
  } else {
    ADP_CHECK_ARG(b, "rows_reduce 2: needs b");
    ADP_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
    ADP_LAUNCH(dot_all_kernel, grid_for((long long)rows * C / 8 / 8), 0, s, (const bf16*)a, (const bf16*)b, (long long)rows * C / 8, out);
  }
  return ADP_OK;
}

extern "C" int adp_softmax_rows(const float* S, int64_t R, int N, float scale, void* P, float* m, float* l, void* stream) {
  ADP_CHECK_ARG(S && P && m && l && R > 0 && N > 0 && N % 4 == 0, "softmax_rows: bad arguments");
  long long blocks = R < (long long)sm_count() * 8 ? R : (long long)sm_count() * 8;
  ADP_LAUNCH(softmax_rows_kernel, (int)blocks, 0, (cudaStream_t)stream, S, (long long)R, N, scale, (bf16*)P, m, l);
  return ADP_OK;
}
extern "C" int adp_softmax_apply(const float* S, int64_t R, int N, float scale, const float* m, const float* l, int by_col, void* P,
                                 void* stream) {
  ADP_CHECK_ARG(S && P && m && l && R > 0 && N > 0 && N % 4 == 0, "softmax_apply: bad arguments");
  ADP_LAUNCH(softmax_apply_kernel, grid_for((long long)R * N / 4), 0, (cudaStream_t)stream, S, (long long)R, N, scale, m, l, by_col, (bf16*)P);
  return ADP_OK;
}
extern "C" int adp_softmax_backward(const void* P, const float* dP, int64_t R, int N, float scale, const float* delta, int by_col,
                                    void* dS, void* stream) {
  ADP_CHECK_ARG(P && dP && delta && dS && R > 0 && N > 0 && N % 4 == 0, "softmax_backward: bad arguments");
  ADP_LAUNCH(softmax_bwd_kernel, grid_for((long long)R * N / 4), 0, (cudaStream_t)stream, (const bf16*)P, dP, (long long)R, N, scale, delta,
             by_col, (bf16*)dS);
  return ADP_OK;
}

extern "C" int adp_gemm_tn_bf16(const void* a, int M, const void* b, int N, float* dw, int ldd, int64_t rows, void* stream) {
  ADP_CHECK_ARG(a && b && dw, "gemm_tn_bf16: null pointer");
  return tc_gemm_tn_full(a, M, b, N, dw, ldd, (long long)rows, (cudaStream_t)stream);
}

extern "C" int adp_depth_head_forward(const void* x, const float* w, const float* bias, float max_depth, int64_t rows, int C, float* y,
                                      void* stream) {
  ADP_CHECK_ARG(x && w && bias && y && rows > 0 && C % 8 == 0, "depth_head_forward: bad arguments");
  ADP_LAUNCH(head_fwd_kernel, grid_for((long long)rows, BT / 32), 0, (cudaStream_t)stream, (const bf16*)x, w, bias, max_depth, (long long)rows, C, y);
  return ADP_OK;
}
extern "C" int adp_depth_head_backward(const void* x, const float* w, const float* bias, float max_depth, const float* dy, int64_t rows,
                                       int C, void* dx, float* dw, float* db, void* stream) {
  ADP_CHECK_ARG(x && w && bias && dy && dx && dw && db && rows > 0 && C % 8 == 0 && C <= 256, "depth_head_backward: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  ADP_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C, s));
  ADP_CUDA(cudaMemsetAsync(db, 0, sizeof(float), s));
  long long blocks = ((long long)rows + (BT / 32) * 32 - 1) / ((BT / 32) * 32);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  ADP_LAUNCH(head_bwd_kernel, (int)(blocks < 1 ? 1 : blocks), (C + 1) * sizeof(float), s, (const bf16*)x, w, bias, max_depth, dy,
             (long long)rows, C, (bf16*)dx, dw, db);
  return ADP_OK;
}

// ---- fused attention epilogues: the score matrix never leaves the GEMM in fp32 -----------------------------------------
namespace {
__global__ void stats_init_kernel(int* __restrict__ m, float* __restrict__ l, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    m[i] = float_to_ordered(-INFINITY);
    l[i] = 0.f;
  }
}
}  // namespace

extern "C" int adp_softmax_stats_init(int* stat_m, float* stat_l, int64_t n, void* stream) {
  ADP_CHECK_ARG(stat_m && stat_l && n > 0, "softmax_stats_init: bad arguments");
  stats_init_kernel<<<adp_cdiv((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(stat_m, stat_l, (long long)n);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

// D[m][n] = sum_k a[m][k] * b[n][k] (bf16, K % 64 == 0, N % 64 == 0) followed by one of the softmax epilogues:
//   mode 1: stat_m[m] = max(stat_m[m], max_n scale*D)   (stat_m holds order-preserving int images of floats)
//   mode 2: stat_l[m] += sum_n exp(scale*D - stat_m[m])
//   mode 3: out[m][n] = exp(scale*D - stat_m[i]) / stat_l[i]
//   mode 4: out[m][n] = scale * pmat[m][n] * (D - delta[i])
// with i = m, or i = n when by_col (D is then the transposed score matrix, statistics stay per query).
extern "C" int adp_gemm_rows_softmax(const void* a, int K, const void* b, int N, int64_t M, int mode, int by_col, float scale,
                                     int* stat_m, float* stat_l, const float* delta, const void* pmat, void* out_bf16,
                                     void* stream) {
  ADP_CHECK_ARG(a && b && mode >= 1 && mode <= 4, "gemm_rows_softmax: bad arguments");
  GemmEpilogue e{mode, by_col, scale, stat_m, stat_l, delta, pmat};
  return tc_gemm_rows(a, K, nullptr, 0, b, 0, mode >= 3 ? out_bf16 : nullptr, N, nullptr, 0, nullptr, (long long)M,
                      (cudaStream_t)stream, &e);
}
