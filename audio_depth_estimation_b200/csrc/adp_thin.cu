// The two "thin" layers of the U-Net, which are HBM-bound rather than tensor-core work
// (E1 = Conv2d(2 -> 64) reads 0.5 MB and writes 4 MB per sample, D1 =
// ConvTranspose2d(128 -> 1) reads 4 MB): models/unetbaseline_model.py:187 (outermost downconv) and
// :196-198 (outermost upconv + bias + ReLU|Sigmoid head).
//
// Shared structure: the one- or two-channel image (network input x, or dL/du of the head) is staged
// per 8x16-pixel tile in shared memory (with its stride-2 halo, zero-filled at the border), a warp
// walks the tile one pixel at a time with its 32 lanes spread over the wide channel dimension, so
// every global access of the wide NHWC tensor is one fully coalesced row segment, the filter taps
// live in registers, and image samples are shared-memory broadcasts.
//   first_conv_fprop / first_conv_wgrad : lanes = output-channel pairs, 16*Cin taps in registers
//   last_convT_dgrad / last_convT_wgrad : lanes = 4 input channels, 16 taps in registers
//   last_convT_col2im                   : head of the tensor-core path: P[pixel][16 taps] -> y
#include "adp_common.cuh"

namespace {

constexpr int TT_H = 8, TT_W = 16;                 // tile of small-grid pixels
constexpr int TR = 2 * TT_H + 2, TC = 2 * TT_W + 2;  // big-grid region incl. halo: 18 x 34
constexpr int TCP = 36;                              // row pitch (floats), multiple of 4
constexpr int TREG = 2 * TR * TCP;                   // two copies: [0] as is, [1] shifted left by 2 columns,
                                                     // so the 4 taps of a kernel row are ONE aligned broadcast LDS.128
constexpr int THIN_THREADS = 256;
constexpr int THIN_WARPS = THIN_THREADS / 32;

__device__ __forceinline__ void load_region(const float* __restrict__ img, int H, int W, int ry0, int rx0,
                                            float* __restrict__ s) {
  for (int idx = threadIdx.x; idx < TR * TC; idx += THIN_THREADS) {
    const int r = idx / TC, c = idx - r * TC;
    const int y = ry0 + r, x = rx0 + c;
    const float v = (y >= 0 && y < H && x >= 0 && x < W) ? img[(size_t)y * W + x] : 0.f;
    s[r * TCP + c] = v;
    if (c >= 2) s[TR * TCP + r * TCP + c - 2] = v;
  }
}
// taps kw = 0..3 of kernel row kh for the tile-local pixel (li, lj): region[2li+kh][2lj .. 2lj+3]
__device__ __forceinline__ float4 taps4(const float* __restrict__ s, int li, int lj, int kh) {
  const int odd = lj & 1;
  return *reinterpret_cast<const float4*>(s + odd * (TR * TCP) + (2 * li + kh) * TCP + 2 * lj - 2 * odd);
}

struct TileIter {
  int tiles_x, tiles_y, ntiles;
  __device__ TileIter(int B, int Hs, int Ws) {
    tiles_x = (Ws + TT_W - 1) / TT_W;
    tiles_y = (Hs + TT_H - 1) / TT_H;
    ntiles = B * tiles_x * tiles_y;
  }
  __device__ void decode(int t, int& b, int& i0, int& j0) const {
    const int tx = t % tiles_x, r = t / tiles_x;
    j0 = tx * TT_W;
    i0 = (r % tiles_y) * TT_H;
    b = r / tiles_y;
  }
};

template <class T> struct Pair;
template <> struct Pair<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Pair<bf16> {
  static __device__ __forceinline__ float2 ld(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
  static __device__ __forceinline__ void st(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

// ------------------------------------------------------------------ first conv, forward
// x NCHW fp32 [B,CIN,H,W], w [N][16][CIN] -> out0/out1 NHWC [B,H/2,W/2,N] (two activations of the same conv)
template <class T, int CIN>
__global__ void __launch_bounds__(THIN_THREADS, 2)
first_conv_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, float slope0, T* __restrict__ out0,
                        float slope1, T* __restrict__ out1, int B, int H, int W, int N) {
  __shared__ __align__(16) float xs[CIN][TREG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ho = H / 2, Wo = W / 2;
  const int n = 2 * lane;
  const bool active = n < N;
  float wr[2][16 * CIN];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int t = 0; t < 16 * CIN; ++t) wr[q][t] = active ? w[(size_t)(n + q) * 16 * CIN + t] : 0.f;
  const TileIter it(B, Ho, Wo);
  for (int tile = blockIdx.x; tile < it.ntiles; tile += gridDim.x) {
    int b, i0, j0;
    it.decode(tile, b, i0, j0);
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
      load_region(x + ((size_t)b * CIN + ci) * H * W, H, W, 2 * i0 - 1, 2 * j0 - 1, xs[ci]);
    __syncthreads();
    for (int p = warp; p < TT_H * TT_W; p += THIN_WARPS) {
      const int li = p / TT_W, lj = p - li * TT_W;
      const int i = i0 + li, j = j0 + lj;
      if (i >= Ho || j >= Wo) continue;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int kh = 0; kh < 4; ++kh)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 t4 = taps4(xs[ci], li, lj, kh);
          const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            a0 = fmaf(tv[kw], wr[0][(kh * 4 + kw) * CIN + ci], a0);
            a1 = fmaf(tv[kw], wr[1][(kh * 4 + kw) * CIN + ci], a1);
          }
        }
      if (active) {
        const size_t o = (((size_t)b * Ho + i) * Wo + j) * N + n;
        Pair<T>::st(out0 + o, lrelu(a0, slope0), lrelu(a1, slope0));
        if (out1) Pair<T>::st(out1 + o, lrelu(a0, slope1), lrelu(a1, slope1));
      }
    }
  }
}

// ------------------------------------------------------------------ first conv, weight gradient
// dw[n][16][CIN] += sum_pixels dy[pix][n] * x[b,ci,2oy-1+kh,2ox-1+kw]
template <class T, int CIN>
__global__ void __launch_bounds__(THIN_THREADS, 2)
first_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int B, int H, int W,
                        int N) {
  __shared__ __align__(16) float xs[CIN][TREG];
  __shared__ float red[64 * 16 * CIN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ho = H / 2, Wo = W / 2;
  const int n = 2 * lane;
  const bool active = n < N;
  for (int i = threadIdx.x; i < 64 * 16 * CIN; i += THIN_THREADS) red[i] = 0.f;
  float acc[2][16 * CIN];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int t = 0; t < 16 * CIN; ++t) acc[q][t] = 0.f;
  const TileIter it(B, Ho, Wo);
  for (int tile = blockIdx.x; tile < it.ntiles; tile += gridDim.x) {
    int b, i0, j0;
    it.decode(tile, b, i0, j0);
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
      load_region(x + ((size_t)b * CIN + ci) * H * W, H, W, 2 * i0 - 1, 2 * j0 - 1, xs[ci]);
    __syncthreads();
    // software pipeline: the next pixel's gradient row is in flight while this pixel's 64 FMAs run
    auto fetch = [&](int p) -> float2 {
      const int li = p / TT_W, lj = p - li * TT_W;
      const int i = i0 + li, j = j0 + lj;
      if (p >= TT_H * TT_W || i >= Ho || j >= Wo || !active) return make_float2(0.f, 0.f);
      return Pair<T>::ld(dy + (((size_t)b * Ho + i) * Wo + j) * N + n);
    };
    float2 g_next = fetch(warp);
    for (int p = warp; p < TT_H * TT_W; p += THIN_WARPS) {
      const int li = p / TT_W, lj = p - li * TT_W;
      const float2 g = g_next;
      g_next = fetch(p + THIN_WARPS);
#pragma unroll
      for (int kh = 0; kh < 4; ++kh)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 t4 = taps4(xs[ci], li, lj, kh);
          const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            acc[0][(kh * 4 + kw) * CIN + ci] = fmaf(g.x, tv[kw], acc[0][(kh * 4 + kw) * CIN + ci]);
            acc[1][(kh * 4 + kw) * CIN + ci] = fmaf(g.y, tv[kw], acc[1][(kh * 4 + kw) * CIN + ci]);
          }
        }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int t = 0; t < 16 * CIN; ++t) atomicAdd(&red[(n + q) * 16 * CIN + t], acc[q][t]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * 16 * CIN; i += THIN_THREADS) atomicAdd(&dw[i], red[i]);
}

// ------------------------------------------------------------------ last transposed conv (Cout = 1), input gradient
// du fp32 [B,1,2Hi,2Wi], w [Ct][16] -> g0 [B,Hi,Wi,C0], g1 [B,Hi,Wi,C1]:  g[pix][c] = sum_taps du[2i-1+kh,2j-1+kw] w[c][kh,kw]
template <class T>
__global__ void __launch_bounds__(THIN_THREADS, 2)
last_convT_dgrad_kernel2(const float* __restrict__ du, const float* __restrict__ w, T* __restrict__ g0, int C0,
                         T* __restrict__ g1, int C1, int B, int Hi, int Wi) {
  __shared__ __align__(16) float ds[TREG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ct = C0 + C1, Ho = 2 * Hi, Wo = 2 * Wi;
  const TileIter it(B, Hi, Wi);
  for (int cb = 0; cb < Ct; cb += 128) {
    const int c = cb + lane * 4;
    const bool active = c < Ct;
    float wr[16][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int t = 0; t < 16; ++t) wr[t][q] = active ? w[(size_t)(c + q) * 16 + t] : 0.f;
    for (int tile = blockIdx.x; tile < it.ntiles; tile += gridDim.x) {
      int b, i0, j0;
      it.decode(tile, b, i0, j0);
      __syncthreads();
      load_region(du + (size_t)b * Ho * Wo, Ho, Wo, 2 * i0 - 1, 2 * j0 - 1, ds);
      __syncthreads();
      for (int p = warp; p < TT_H * TT_W; p += THIN_WARPS) {
        const int li = p / TT_W, lj = p - li * TT_W;
        const int i = i0 + li, j = j0 + lj;
        if (i >= Hi || j >= Wi) continue;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kh = 0; kh < 4; ++kh) {
          const float4 t4 = taps4(ds, li, lj, kh);
          const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            const float d = tv[kw];
            a.x = fmaf(d, wr[kh * 4 + kw][0], a.x); a.y = fmaf(d, wr[kh * 4 + kw][1], a.y);
            a.z = fmaf(d, wr[kh * 4 + kw][2], a.z); a.w = fmaf(d, wr[kh * 4 + kw][3], a.w);
          }
        }
        if (active) {
          const size_t pix = ((size_t)b * Hi + i) * Wi + j;
          if (c < C0) st4(g0 + pix * C0 + c, a);
          else st4(g1 + pix * C1 + (c - C0), a);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ last transposed conv, weight gradient
// dw[c][16] += sum_pixels x[pix][c] * du[2i-1+kh,2j-1+kw];  x = (x0 | x1) NHWC
template <class T>
__global__ void __launch_bounds__(THIN_THREADS, 2)
last_convT_wgrad_kernel2(const T* __restrict__ x0, int C0, const T* __restrict__ x1, int C1, const float* __restrict__ du,
                         float* __restrict__ dw, int B, int Hi, int Wi, int cb) {
  __shared__ __align__(16) float ds[TREG];
  __shared__ float red[16][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ct = C0 + C1, Ho = 2 * Hi, Wo = 2 * Wi;
  const int c = cb + lane * 4;
  const bool active = c < Ct;
  for (int i = threadIdx.x; i < 16 * 128; i += THIN_THREADS) (&red[0][0])[i] = 0.f;
  float acc[16][4];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  const TileIter it(B, Hi, Wi);
  for (int tile = blockIdx.x; tile < it.ntiles; tile += gridDim.x) {
    int b, i0, j0;
    it.decode(tile, b, i0, j0);
    __syncthreads();
    load_region(du + (size_t)b * Ho * Wo, Ho, Wo, 2 * i0 - 1, 2 * j0 - 1, ds);
    __syncthreads();
    auto fetch = [&](int p) -> float4 {
      const int li = p / TT_W, lj = p - li * TT_W;
      const int i = i0 + li, j = j0 + lj;
      if (p >= TT_H * TT_W || i >= Hi || j >= Wi || !active) return make_float4(0.f, 0.f, 0.f, 0.f);
      const size_t pix = ((size_t)b * Hi + i) * Wi + j;
      return c < C0 ? ld4(x0 + pix * C0 + c) : ld4(x1 + pix * C1 + (c - C0));
    };
    float4 v_next = fetch(warp);
    for (int p = warp; p < TT_H * TT_W; p += THIN_WARPS) {
      const int li = p / TT_W, lj = p - li * TT_W;
      const float4 v = v_next;
      v_next = fetch(p + THIN_WARPS);
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        const float4 t4 = taps4(ds, li, lj, kh);
        const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const float d = tv[kw];
          float* a = acc[kh * 4 + kw];
          a[0] = fmaf(d, v.x, a[0]); a[1] = fmaf(d, v.y, a[1]); a[2] = fmaf(d, v.z, a[2]); a[3] = fmaf(d, v.w, a[3]);
        }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(&red[t][lane * 4 + q], acc[t][q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * 128; i += THIN_THREADS) {
    const int t = i >> 7, cl = i & 127;
    if (cb + cl < Ct) atomicAdd(&dw[(size_t)(cb + cl) * 16 + t], red[t][cl]);
  }
}

// ------------------------------------------------------------------ head of the tensor-core forward path
// P fp32 [B,Hi,Wi,16] (per input pixel: sum_c x[c] w[c][tap]) -> y[b,2i+a,2j+bb] = act(bias + sum of the 4 taps that land there)
__global__ void __launch_bounds__(256)
last_convT_col2im_kernel(const float* __restrict__ P, const float* __restrict__ bias, int final_sigmoid,
                         float* __restrict__ y, int B, int Hi, int Wi) {
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const long long total = (long long)B * Ho * Wo;
  const float bv = bias ? bias[0] : 0.f;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(o % Wo);
    const long long r = o / Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    const int a = oy & 1, bb = ox & 1, i = oy >> 1, j = ox >> 1;
    float u = bv;
#pragma unroll
    for (int th = 0; th < 2; ++th) {
      const int iy = i + a - 1 + th, kh = 3 - a - 2 * th;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int tw = 0; tw < 2; ++tw) {
        const int ix = j + bb - 1 + tw, kw = 3 - bb - 2 * tw;
        if (ix < 0 || ix >= Wi) continue;
        u += P[(((size_t)b * Hi + iy) * Wi + ix) * 16 + kh * 4 + kw];
      }
    }
    y[o] = final_sigmoid ? 1.f / (1.f + expf(-u)) : fmaxf(u, 0.f);
  }
}

// ------------------------------------------------------------------ tensor-core route for the thin layers
// Patch matrices (im2col rows padded to 64 bf16 = one 128-byte swizzled TMA row):
//   mode 0: xp[pix][t], t = (kh*4+kw)*Cin + ci < 16*Cin, = x[b,ci,2oy-1+kh,2ox-1+kw]   (x NCHW fp32, Cin planes)
//   mode 1: dp[pix][t], t = kh*4+kw < 16,              = du[b,2i-1+kh,2j-1+kw]          (same with Cin = 1)
// One thread writes one 16-byte chunk (8 values).
// SPLIT (needs 16*CIN <= 32): columns 0..31 hold bf16(x) and columns 32..63 hold bf16(x - bf16(x)), so that the
// network input keeps ~16 mantissa bits through the bf16 tensor-core GEMM (the weight rows are duplicated).
template <int CIN, bool SPLIT>
__global__ void __launch_bounds__(256)
patch_rows_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int H, int W, int lw, int lh) {
  // Wo = 1 << lw, Ho = 1 << lh (the tensor-core path needs power-of-two grids); 32-bit indices
  const int Wo = 1 << lw, Ho = 1 << lh;
  const unsigned total = (unsigned)B * Ho * Wo * 8u;
  constexpr int K = 16 * CIN;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int q8 = (int)(idx & 7u);
    const bool lo_half = SPLIT && q8 >= 4;
    const int q = SPLIT ? (q8 & 3) : q8;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (q * 8 < K) {
      const unsigned pix = idx >> 3;
      const int ox = (int)(pix & (Wo - 1)), oy = (int)((pix >> lw) & (Ho - 1)), b = (int)(pix >> (lw + lh));
      float r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int t = q * 8 + e;                 // compile-time after unrolling except for q
        const int tap = t / CIN, ci = t % CIN;
        const int y = 2 * oy - 1 + (tap >> 2), x = 2 * ox - 1 + (tap & 3);
        r[e] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (((size_t)b * CIN + ci) * H + y) * W + x) : 0.f;
        if (lo_half) r[e] -= __bfloat162float(__float2bfloat16_rn(r[e]));
      }
      u.x = pack_bf16x2(r[0], r[1]); u.y = pack_bf16x2(r[2], r[3]);
      u.z = pack_bf16x2(r[4], r[5]); u.w = pack_bf16x2(r[6], r[7]);
    }
    *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = u;
  }
}

// src fp32 [R][K] -> dst bf16 [R][64] zero padded (K <= 64); dup: columns 32..32+K-1 repeat columns 0..K-1 (K <= 32)
__global__ void pad_rows_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int R, int K, int dup) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 64) return;
  const int r = i >> 6, k = i & 63;
  const int kk = (dup && k >= 32) ? k - 32 : k;
  dst[i] = __float2bfloat16_rn(kk < K ? src[(size_t)r * K + kk] : 0.f);
}

// D fp32 [128][NT] from the transposed-operand GEMM -> dw
//   mode 0 (last convT): dw[c][t]   = D[c][t],                 c < 128, t < 16    (ldd = 64)
//   mode 1 (first conv): dw[n][t]   = D[n][t] + D[64+n][64+t], n < 64,  t < K     (ldd = 128; pixel pairs folded)
//   mode 2 (first conv, hi/lo split input): as mode 1 plus the low-order halves D[n][32+t] + D[64+n][96+t]
__global__ void fold_thin_wgrad_kernel(const float* __restrict__ D, float* __restrict__ dw, int mode, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == 0) {
    if (i < 128 * 16) dw[i] = D[(i >> 4) * 64 + (i & 15)];
  } else {
    if (i < 64 * K) {
      const int n = i / K, t = i - n * K;
      float v = D[n * 128 + t] + D[(64 + n) * 128 + 64 + t];
      if (mode == 2) v += D[n * 128 + 32 + t] + D[(64 + n) * 128 + 96 + t];
      dw[i] = v;
    }
  }
}


// ------------------------------------------------------------------ first-level centring (adp_unet.cu: use_center)
// xsum[ci] += sum over the batch of plane ci; x NCHW fp32 [B][Cin][HW]; grid (blocks, B * Cin)
__global__ void __launch_bounds__(256)
center_input_sums_kernel(const float* __restrict__ x, int Cin, long long HW, double* __restrict__ xsum) {
  const int plane = blockIdx.y, ci = plane % Cin;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)plane * HW);
  const long long n4 = HW / 4;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    acc += (v.x + v.y) + (v.z + v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = 4 * n4 + threadIdx.x; i < HW; i += blockDim.x) acc += x[(size_t)plane * HW + i];
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[w];
    atomicAdd(&xsum[ci], t);
  }
}

// blocks [0, N2): T[n2] = sum_{tap,c} w2b[n2][tap][c] * m[c];  blocks >= N2: border ring of a0pad = -m.
// Every block derives m itself (64 x 16*Cin multiply-adds): m[n] = bf16(LeakyReLU_0.2(sum_t w1[n][t] * mean_x[t % Cin])),
// rounded to bf16 so that -m is stored exactly in the ring.  Block 0 publishes m.
__global__ void __launch_bounds__(256)
center_tables_kernel(const double* __restrict__ xsum, double inv_count, const float* __restrict__ w1, int Cin,
                     const bf16* __restrict__ w2b, int N2, float* __restrict__ m_out, float* __restrict__ T,
                     bf16* __restrict__ a0pad, int B, int H, int W) {
  __shared__ float m_s[64];
  __shared__ float red[8];
  if (threadIdx.x < 64) {
    const int n = threadIdx.x, K = 16 * Cin;
    float acc = 0.f;
    for (int t = 0; t < K; ++t) acc = fmaf(w1[n * K + t], (float)(xsum[t % Cin] * inv_count), acc);
    const float mv = __bfloat162float(__float2bfloat16_rn(lrelu(acc, 0.2f)));
    m_s[n] = mv;
    if (blockIdx.x == 0) m_out[n] = mv;
  }
  __syncthreads();
  if ((int)blockIdx.x < N2) {
    const bf16* row = w2b + (size_t)blockIdx.x * 16 * 64;
    float acc = 0.f;
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) acc = fmaf(__bfloat162float(row[i]), m_s[i & 63], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      T[blockIdx.x] = t;
    }
    return;
  }
  // ring: 2 (W+2) + 2 H pixels per sample, 8 chunks of 8 channels per pixel
  const int Hp = H + 2, Wp = W + 2, ring = 2 * Wp + 2 * H;
  const long long total = (long long)B * ring * 8;
  const int rblocks = gridDim.x - N2;
  for (long long idx = (long long)(blockIdx.x - N2) * blockDim.x + threadIdx.x; idx < total; idx += (long long)rblocks * blockDim.x) {
    const int ch8 = (int)(idx & 7);
    const long long rp = idx >> 3;
    const int r = (int)(rp % ring), b = (int)(rp / ring);
    int yy, xx;
    if (r < Wp) { yy = 0; xx = r; }
    else if (r < 2 * Wp) { yy = Hp - 1; xx = r - Wp; }
    else { const int k = r - 2 * Wp; yy = 1 + (k >> 1); xx = (k & 1) ? Wp - 1 : 0; }
    uint4 u;
    u.x = pack_bf16x2(-m_s[ch8 * 8 + 0], -m_s[ch8 * 8 + 1]); u.y = pack_bf16x2(-m_s[ch8 * 8 + 2], -m_s[ch8 * 8 + 3]);
    u.z = pack_bf16x2(-m_s[ch8 * 8 + 4], -m_s[ch8 * 8 + 5]); u.w = pack_bf16x2(-m_s[ch8 * 8 + 6], -m_s[ch8 * 8 + 7]);
    *reinterpret_cast<uint4*>(a0pad + (((size_t)b * Hp + yy) * Wp + xx) * 64 + ch8 * 8) = u;
  }
}

__global__ void center_wgrad_fix_kernel(float* __restrict__ dw, const float* __restrict__ m, const float* __restrict__ scale,
                                        const double* __restrict__ gsum, int N2, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N2 * 16 * C) return;
  const int c = (int)(i % C), n = (int)(i / (16LL * C));
  dw[i] += m[c] * scale[n] * (float)gsum[n];
}

int tile_grid(int B, int Hs, int Ws, int per_sm) {
  long long tiles = (long long)B * ((Hs + TT_H - 1) / TT_H) * ((Ws + TT_W - 1) / TT_W);
  long long cap = (long long)adp::sm_count() * per_sm;
  return (int)(tiles < cap ? tiles : cap);
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

#define ADP_DISPATCH_CIN(cin, ...)                                             \
  switch (cin) {                                                               \
    case 1: { constexpr int CIN = 1; __VA_ARGS__ } break;                      \
    case 2: { constexpr int CIN = 2; __VA_ARGS__ } break;                      \
    case 3: { constexpr int CIN = 3; __VA_ARGS__ } break;                      \
    case 4: { constexpr int CIN = 4; __VA_ARGS__ } break;                      \
    default: adp_set_error("thin conv: Cin %d unsupported", cin); return ADP_ERR_UNSUPPORTED; \
  }

bool thin_first_supported(int Cin, int N) { return Cin >= 1 && Cin <= 4 && N <= 64 && N % 2 == 0; }

int thin_first_conv_fprop(int dtype, const float* x, const float* w, float slope0, void* out0, float slope1, void* out1,
                          int B, int H, int W, int Cin, int N, cudaStream_t s) {
  const int grid = tile_grid(B, H / 2, W / 2, 8);
  ADP_DISPATCH_T(dtype, ADP_DISPATCH_CIN(Cin, {
    first_conv_fprop_kernel<T, CIN><<<grid, THIN_THREADS, 0, s>>>(x, w, slope0, (T*)out0, slope1, (T*)out1, B, H, W, N);
  }))
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int thin_first_conv_wgrad(int dtype, const float* x, const void* dy, float* dw, int B, int H, int W, int Cin, int N,
                          cudaStream_t s) {
  const int grid = tile_grid(B, H / 2, W / 2, 2);
  ADP_DISPATCH_T(dtype, ADP_DISPATCH_CIN(Cin, {
    first_conv_wgrad_kernel<T, CIN><<<grid, THIN_THREADS, 0, s>>>(x, (const T*)dy, dw, B, H, W, N);
  }))
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int thin_last_convT_dgrad(int dtype, const float* du, const float* w, void* g0, int C0, void* g1, int C1, int B, int Hi,
                          int Wi, cudaStream_t s) {
  const int grid = tile_grid(B, Hi, Wi, 8);
  ADP_DISPATCH_T(dtype, {
    last_convT_dgrad_kernel2<T><<<grid, THIN_THREADS, 0, s>>>(du, w, (T*)g0, C0, (T*)g1, C1, B, Hi, Wi);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int thin_last_convT_wgrad(int dtype, const void* x0, int C0, const void* x1, int C1, const float* du, float* dw, int B,
                          int Hi, int Wi, cudaStream_t s) {
  const int grid = tile_grid(B, Hi, Wi, 2);
  for (int cb = 0; cb < C0 + C1; cb += 128) {
    ADP_DISPATCH_T(dtype, {
      last_convT_wgrad_kernel2<T><<<grid, THIN_THREADS, 0, s>>>((const T*)x0, C0, (const T*)x1, C1, du, dw, B, Hi, Wi, cb);
    })
    ADP_LAUNCH_CHECK();
  }
  return ADP_OK;
}

int thin_patch_rows(const float* img, void* out, int B, int Cin, int H, int W, int split, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2;
  int lw = 0, lh = 0;
  while ((1 << lw) < Wo) ++lw;
  while ((1 << lh) < Ho) ++lh;
  ADP_CHECK_ARG((1 << lw) == Wo && (1 << lh) == Ho && (long long)B * Ho * Wo * 8 < (1LL << 31),
                "patch_rows: grid must be a power of two and fit 32-bit indexing");
  long long total = (long long)B * Ho * Wo * 8;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ADP_CHECK_ARG(!split || 16 * Cin <= 32, "patch_rows: hi/lo split needs 16*Cin <= 32");
  switch (Cin * 2 + (split ? 1 : 0)) {
    case 2: patch_rows_kernel<1, false><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    case 3: patch_rows_kernel<1, true><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    case 4: patch_rows_kernel<2, false><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    case 5: patch_rows_kernel<2, true><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    case 6: patch_rows_kernel<3, false><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    case 8: patch_rows_kernel<4, false><<<(int)blocks, 256, 0, s>>>(img, (bf16*)out, B, H, W, lw, lh); break;
    default: adp_set_error("patch_rows: Cin %d unsupported", Cin); return ADP_ERR_UNSUPPORTED;
  }
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}


int center_input_sums(const float* x, int B, int Cin, long long HW, double* xsum, cudaStream_t s) {
  ADP_CHECK_ARG(Cin >= 1 && Cin <= 16 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && HW % 4 == 0,
                "center_input_sums: unsupported input");
  long long bx = (HW / 4 + 255) / 256;
  if (bx > 8) bx = 8;
  if (bx < 1) bx = 1;
  center_input_sums_kernel<<<dim3((unsigned)bx, (unsigned)(B * Cin)), 256, 0, s>>>(x, Cin, HW, xsum);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int center_tables(const double* xsum, double inv_count, const float* w1, int Cin, const void* w2b, int N2, float* m,
                  float* T, void* a0pad, int B, int H, int W, cudaStream_t s) {
  const long long ring_items = (long long)B * (2 * (W + 2) + 2 * H) * 8;
  long long rblocks = (ring_items + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (rblocks > cap) rblocks = cap;
  center_tables_kernel<<<(unsigned)(N2 + rblocks), 256, 0, s>>>(xsum, inv_count, w1, Cin, (const bf16*)w2b, N2, m, T,
                                                                 (bf16*)a0pad, B, H, W);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int center_wgrad_fix(float* dw, const float* m, const float* scale, const double* gsum, int N2, int C, cudaStream_t s) {
  const long long n = (long long)N2 * 16 * C;
  center_wgrad_fix_kernel<<<adp_cdiv(n, 256), 256, 0, s>>>(dw, m, scale, gsum, N2, C);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int thin_pad_rows(const float* src, void* dst, int R, int K, int dup, cudaStream_t s) {
  ADP_CHECK_ARG(K <= 64 && (!dup || K <= 32), "pad_rows: K too large");
  pad_rows_kernel<<<adp_cdiv((long long)R * 64, 256), 256, 0, s>>>(src, (bf16*)dst, R, K, dup);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int thin_fold_wgrad(const float* D, float* dw, int mode, int K, cudaStream_t s) {
  const int n = mode == 0 ? 128 * 16 : 64 * K;
  fold_thin_wgrad_kernel<<<adp_cdiv(n, 256), 256, 0, s>>>(D, dw, mode, K);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int last_convT_col2im(const float* P, const float* bias, int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s) {
  long long total = (long long)B * 4 * Hi * Wi;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  last_convT_col2im_kernel<<<(int)blocks, 256, 0, s>>>(P, bias, final_sigmoid, y, B, Hi, Wi);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

}  // namespace adp
