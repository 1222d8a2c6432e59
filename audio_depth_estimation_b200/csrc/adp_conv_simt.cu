// fp32-accumulate SIMT convolutions: the full-precision (ADP_F32) path of the U-Net and the
// layers that are HBM-bound rather than tensor-core work (first conv, Cin = 2; last
// transposed conv, Cout = 1).
//
// Replaces cuDNN/oneDNN behind nn.Conv2d(k4,s2,p1) (models/unetbaseline_model.py:187) and
// nn.ConvTranspose2d(k4,s2,p1) (:196,:209,:218), forward, dgrad and wgrad, as three implicit
// GEMM families over NHWC tensors:
//   F1 "gather"  : y[b,oy,ox,n]      = sum_{kh,kw,c} x[b,2oy-1+kh,2ox-1+kw,c] * w[n][kh,kw][c]
//                  (Conv2d fprop; ConvTranspose2d dgrad with the transposed-conv weight)
//   F2 "parity"  : y[b,2i+a,2j+bb,n] = sum_{th,tw,c} x[b,i+a-1+th,j+bb-1+tw,c] * w[c][kh,kw][n],
//                  kh = 3-a-2th, kw = 3-bb-2tw   (ConvTranspose2d fprop; Conv2d dgrad)
//   F3 "wgrad"   : dw[m][kh,kw][n]  += sum_{b,i,j} S[b,i,j,m] * G[b,2i-1+kh,2j-1+kw,n]
// One 64x64x16 register-tiled kernel serves all three through small "problem" functors.
#include "adp_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS_PAD = 4;

template <class P>
__global__ void __launch_bounds__(256) igemm_kernel(const P p) {
  __shared__ __align__(16) float As[BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const typename P::Ctx ctx = p.make_ctx(blockIdx.z);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = ctx.kbeg; k0 < ctx.kend; k0 += BK) {
    if (P::A_VEC_K) {
      const int m = tid >> 2, kk = (tid & 3) * 4;
      float4 v = p.loadA(ctx, m0 + m, k0 + kk);
      As[kk + 0][m] = v.x; As[kk + 1][m] = v.y; As[kk + 2][m] = v.z; As[kk + 3][m] = v.w;
    } else {
      const int kk = tid >> 4, m = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&As[kk][m]) = p.loadA(ctx, m0 + m, k0 + kk);
    }
    if (P::B_VEC_K) {
      const int n = tid >> 2, kk = (tid & 3) * 4;
      float4 v = p.loadB(ctx, k0 + kk, n0 + n);
      Bs[kk + 0][n] = v.x; Bs[kk + 1][n] = v.y; Bs[kk + 2][n] = v.z; Bs[kk + 3][n] = v.w;
    } else {
      const int kk = tid >> 4, n = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&Bs[kk][n]) = p.loadB(ctx, k0 + kk, n0 + n);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    p.storeC(ctx, m0 + ty * 4 + i, n0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
}

#define Z4 make_float4(0.f, 0.f, 0.f, 0.f)

struct CtxRange {
  int kbeg, kend;
  int a, bb;   // F2 parity
  int kh, kw;  // F3 tap
};

// ------------------------------------------------------------------ F1
template <class T>
struct GatherConv {
  static constexpr bool A_VEC_K = true, B_VEC_K = true;
  using Ctx = CtxRange;
  const T* x; const float* w; T* y0; T* y1;
  int N0, N1, B, Hi, Wi, C, Ho, Wo, M, N, K;
  __device__ Ctx make_ctx(int) const { return Ctx{0, K, 0, 0, 0, 0}; }
  __device__ float4 loadA(const Ctx&, int m, int k) const {
    if (m >= M || k >= K) return Z4;
    int ox = m % Wo, t = m / Wo, oy = t % Ho, b = t / Ho;
    int tap = k / C, c = k - tap * C;
    int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
    if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return Z4;
    return ld4(x + (((size_t)b * Hi + iy) * Wi + ix) * C + c);
  }
  __device__ float4 loadB(const Ctx&, int k, int n) const {
    if (n >= N || k >= K) return Z4;
    return ld4(w + (size_t)n * K + k);
  }
  __device__ void storeC(const Ctx&, int m, int n, float4 v) const {
    if (m >= M || n >= N) return;
    if (n < N0) st4(y0 + (size_t)m * N0 + n, v);
    else st4(y1 + (size_t)m * N1 + (n - N0), v);
  }
};

// first conv: NCHW fp32 input with a handful of channels, two activated outputs
template <class T>
struct FirstConv {
  static constexpr bool A_VEC_K = true, B_VEC_K = true;
  using Ctx = CtxRange;
  const float* x; const float* w; T* out0; T* out1; float slope0, slope1;
  int B, Hi, Wi, C, Ho, Wo, M, N, K;
  __device__ Ctx make_ctx(int) const { return Ctx{0, K, 0, 0, 0, 0}; }
  __device__ float one(int b, int oy, int ox, int k) const {
    if (k >= K) return 0.f;
    int tap = k / C, c = k - tap * C;
    int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
    if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return 0.f;
    return x[(((size_t)b * C + c) * Hi + iy) * Wi + ix];
  }
  __device__ float4 loadA(const Ctx&, int m, int k) const {
    if (m >= M) return Z4;
    int ox = m % Wo, t = m / Wo, oy = t % Ho, b = t / Ho;
    return make_float4(one(b, oy, ox, k), one(b, oy, ox, k + 1), one(b, oy, ox, k + 2), one(b, oy, ox, k + 3));
  }
  __device__ float4 loadB(const Ctx&, int k, int n) const {
    if (n >= N || k >= K) return Z4;
    return ld4(w + (size_t)n * K + k);
  }
  __device__ void storeC(const Ctx&, int m, int n, float4 v) const {
    if (m >= M || n >= N) return;
    st4(out0 + (size_t)m * N + n, make_float4(lrelu(v.x, slope0), lrelu(v.y, slope0), lrelu(v.z, slope0), lrelu(v.w, slope0)));
    if (out1)
      st4(out1 + (size_t)m * N + n, make_float4(lrelu(v.x, slope1), lrelu(v.y, slope1), lrelu(v.z, slope1), lrelu(v.w, slope1)));
  }
};

// ------------------------------------------------------------------ F2
template <class T>
struct ParityConvT {
  static constexpr bool A_VEC_K = true, B_VEC_K = false;
  using Ctx = CtxRange;
  const T* x0; const T* x1; const float* w; T* y;
  int C0, C1, Ct, B, Hi, Wi, M, N, K;
  __device__ Ctx make_ctx(int z) const { return Ctx{0, K, z >> 1, z & 1, 0, 0}; }
  __device__ float4 loadA(const Ctx& cx, int m, int k) const {
    if (m >= M || k >= K) return Z4;
    int j = m % Wi, t = m / Wi, i = t % Hi, b = t / Hi;
    int tap2 = k / Ct, c = k - tap2 * Ct;
    int iy = i + cx.a - 1 + (tap2 >> 1), ix = j + cx.bb - 1 + (tap2 & 1);
    if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return Z4;
    size_t pix = ((size_t)b * Hi + iy) * Wi + ix;
    return c < C0 ? ld4(x0 + pix * C0 + c) : ld4(x1 + pix * C1 + (c - C0));
  }
  __device__ float4 loadB(const Ctx& cx, int k, int n) const {
    if (n >= N || k >= K) return Z4;
    int tap2 = k / Ct, c = k - tap2 * Ct;
    int kh = 3 - cx.a - 2 * (tap2 >> 1), kw = 3 - cx.bb - 2 * (tap2 & 1);
    return ld4(w + ((size_t)c * 16 + kh * 4 + kw) * N + n);
  }
  __device__ void storeC(const Ctx& cx, int m, int n, float4 v) const {
    if (m >= M || n >= N) return;
    int j = m % Wi, t = m / Wi, i = t % Hi, b = t / Hi;
    st4(y + ((((size_t)b * 2 * Hi) + 2 * i + cx.a) * (2 * Wi) + 2 * j + cx.bb) * N + n, v);
  }
};

// ------------------------------------------------------------------ F3
template <class T>
struct Wgrad {
  static constexpr bool A_VEC_K = false, B_VEC_K = false;
  using Ctx = CtxRange;
  const T* s0; const T* s1; const T* g; float* dw;
  int M0, M1, M, N, B, Hs, Ws, K, chunk;
  __device__ Ctx make_ctx(int z) const {
    int tap = z & 15, sk = z >> 4;
    int kb = sk * chunk;
    return Ctx{kb, min(kb + chunk, K), 0, 0, tap >> 2, tap & 3};
  }
  __device__ float4 loadA(const Ctx& cx, int m, int k) const {
    if (m >= M || k >= cx.kend) return Z4;
    return m < M0 ? ld4(s0 + (size_t)k * M0 + m) : ld4(s1 + (size_t)k * M1 + (m - M0));
  }
  __device__ float4 loadB(const Ctx& cx, int k, int n) const {
    if (n >= N || k >= cx.kend) return Z4;
    int j = k % Ws, t = k / Ws, i = t % Hs, b = t / Hs;
    int gy = 2 * i - 1 + cx.kh, gx = 2 * j - 1 + cx.kw;
    if (gy < 0 || gy >= 2 * Hs || gx < 0 || gx >= 2 * Ws) return Z4;
    return ld4(g + (((size_t)b * 2 * Hs + gy) * (2 * Ws) + gx) * N + n);
  }
  __device__ void storeC(const Ctx& cx, int m, int n, float4 v) const {
    if (m >= M || n >= N) return;
    float* d = dw + ((size_t)m * 16 + cx.kh * 4 + cx.kw) * N + n;
    atomicAdd(d + 0, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
  }
};

// first-conv wgrad: dw[n][tap*Cin+ci] += sum_pixels dy[p][n] * x[b,ci,2oy-1+kh,2ox-1+kw]
template <class T>
struct FirstWgrad {
  static constexpr bool A_VEC_K = false, B_VEC_K = false;
  using Ctx = CtxRange;
  const float* x; const T* dy; float* dw;
  int B, Hi, Wi, C, Ho, Wo, M /*=N_out*/, N /*=16*C*/, K /*pixels*/, chunk;
  __device__ Ctx make_ctx(int z) const {
    int kb = z * chunk;
    return Ctx{kb, min(kb + chunk, K), 0, 0, 0, 0};
  }
  __device__ float4 loadA(const Ctx& cx, int m, int k) const {
    if (m >= M || k >= cx.kend) return Z4;
    return ld4(dy + (size_t)k * M + m);
  }
  __device__ float one(int b, int oy, int ox, int q) const {
    if (q >= N) return 0.f;
    int tap = q / C, c = q - tap * C;
    int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
    if (iy < 0 || iy >= Hi || ix < 0 || ix >= Wi) return 0.f;
    return x[(((size_t)b * C + c) * Hi + iy) * Wi + ix];
  }
  __device__ float4 loadB(const Ctx& cx, int k, int n) const {
    if (k >= cx.kend || n >= N) return Z4;
    int ox = k % Wo, t = k / Wo, oy = t % Ho, b = t / Ho;
    return make_float4(one(b, oy, ox, n), one(b, oy, ox, n + 1), one(b, oy, ox, n + 2), one(b, oy, ox, n + 3));
  }
  __device__ void storeC(const Ctx&, int m, int n, float4 v) const {
    if (m >= M) return;
    float* d = dw + (size_t)m * N + n;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n + q < N) atomicAdd(d + q, vv[q]);
  }
};

int pick_split(long long tiles, int K) {
  long long want = ((long long)adp::sm_count() * 4 + tiles - 1) / tiles;
  long long maxs = K / (BK * 8);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 4096) want = 4096;
  return (int)want;
}

// ------------------------------------------------------------------ last transposed conv (Cout = 1)
// One warp per input pixel (b,i,j); lane l owns channels {4l + 128q}.  Weights staged as ws[tap][c].
constexpr int LAST_THREADS = 256;

template <class T>
__device__ __forceinline__ float4 ld_cat(const T* x0, int C0, const T* x1, int C1, size_t pix, int c) {
  return c < C0 ? ld4(x0 + pix * C0 + c) : ld4(x1 + pix * C1 + (c - C0));
}

template <class T>
__global__ void __launch_bounds__(LAST_THREADS)
last_convT_fprop_kernel(const T* __restrict__ x0, int C0, const T* __restrict__ x1, int C1,
                        const float* __restrict__ w, const float* __restrict__ bias, int final_sigmoid,
                        float* __restrict__ y, int B, int Hi, int Wi) {
  extern __shared__ __align__(16) float ws[];  // [16][Ct]
  const int Ct = C0 + C1;
  for (int i = threadIdx.x; i < 16 * Ct; i += LAST_THREADS) {
    int tap = i / Ct, c = i - tap * Ct;
    ws[i] = w[(size_t)c * 16 + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * LAST_THREADS + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * LAST_THREADS) >> 5;
  const long long npix = (long long)B * Hi * Wi;
  const float bv = bias ? bias[0] : 0.f;
  for (long long p = warp; p < npix; p += nwarps) {
    const int j = (int)(p % Wi);
    const long long t = p / Wi;
    const int i = (int)(t % Hi), b = (int)(t / Hi);
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int c = lane * 4; c < Ct; c += 128) {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int iy = i + dy;
        if (iy < 0 || iy >= Hi) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int ix = j + dx;
          if (ix < 0 || ix >= Wi) continue;
          float4 v = ld_cat(x0, C0, x1, C1, ((size_t)b * Hi + iy) * Wi + ix, c);
          // input offset dy contributes to output parity a with dy = a - 1 + th  ->  kh = 3 - a - 2th
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int th = dy - a + 1;
            if (th < 0 || th > 1) continue;
            const int kh = 3 - a - 2 * th;
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const int tw = dx - bb + 1;
              if (tw < 0 || tw > 1) continue;
              const int kw = 3 - bb - 2 * tw;
              float4 wv = *reinterpret_cast<const float4*>(&ws[(kh * 4 + kw) * Ct + c]);
              acc[a][bb] += v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) acc[a][bb] = warp_sum(acc[a][bb]);
    if (lane < 4) {
      const int a = lane >> 1, bb = lane & 1;
      float u = (a ? (bb ? acc[1][1] : acc[1][0]) : (bb ? acc[0][1] : acc[0][0])) + bv;
      float r = final_sigmoid ? 1.f / (1.f + expf(-u)) : fmaxf(u, 0.f);
      y[((size_t)b * 2 * Hi + 2 * i + a) * (2 * Wi) + 2 * j + bb] = r;
    }
  }
}

int pix_grid(long long npix) {
  long long blocks = (npix + (LAST_THREADS / 32) * 4 - 1) / ((LAST_THREADS / 32) * 4);
  long long cap = (long long)adp::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
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

int simt_gather_conv(int dtype, const void* x, const float* w, void* y0, int N0, void* y1, int N1,
                     int B, int Hi, int Wi, int C, cudaStream_t s) {
  ADP_CHECK_ARG(C % 4 == 0 && N0 % 4 == 0 && N1 % 4 == 0 && Hi % 2 == 0 && Wi % 2 == 0,
                "gather_conv: C, N0, N1 must be multiples of 4 and H, W even");
  ADP_DISPATCH_T(dtype, {
    GatherConv<T> p;
    p.x = (const T*)x; p.w = w; p.y0 = (T*)y0; p.y1 = (T*)y1; p.N0 = N0; p.N1 = N1;
    p.B = B; p.Hi = Hi; p.Wi = Wi; p.C = C; p.Ho = Hi / 2; p.Wo = Wi / 2;
    p.M = B * p.Ho * p.Wo; p.N = N0 + N1; p.K = 16 * C;
    dim3 grid(adp_cdiv(p.M, BM), adp_cdiv(p.N, BN), 1);
    igemm_kernel<<<grid, 256, 0, s>>>(p);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int first_conv_fprop(int dtype, const float* x, const float* w, float slope0, void* out0, float slope1,
                     void* out1, int B, int H, int W, int Cin, int N, cudaStream_t s) {
  ADP_CHECK_ARG(N % 4 == 0 && (16 * Cin) % 4 == 0 && H % 2 == 0 && W % 2 == 0, "first_conv: bad shape");
  if (thin_first_supported(Cin, N))
    return thin_first_conv_fprop(dtype, x, w, slope0, out0, slope1, out1, B, H, W, Cin, N, s);
  ADP_DISPATCH_T(dtype, {
    FirstConv<T> p;
    p.x = x; p.w = w; p.out0 = (T*)out0; p.out1 = (T*)out1; p.slope0 = slope0; p.slope1 = slope1;
    p.B = B; p.Hi = H; p.Wi = W; p.C = Cin; p.Ho = H / 2; p.Wo = W / 2;
    p.M = B * p.Ho * p.Wo; p.N = N; p.K = 16 * Cin;
    dim3 grid(adp_cdiv(p.M, BM), adp_cdiv(p.N, BN), 1);
    igemm_kernel<<<grid, 256, 0, s>>>(p);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int simt_parity_convT(int dtype, const void* x0, int C0, const void* x1, int C1, const float* w, void* y,
                      int B, int Hi, int Wi, int N, cudaStream_t s) {
  ADP_CHECK_ARG(C0 % 4 == 0 && C1 % 4 == 0 && N % 4 == 0, "parity_convT: C0, C1, N must be multiples of 4");
  ADP_DISPATCH_T(dtype, {
    ParityConvT<T> p;
    p.x0 = (const T*)x0; p.x1 = (const T*)x1; p.w = w; p.y = (T*)y; p.C0 = C0; p.C1 = C1; p.Ct = C0 + C1;
    p.B = B; p.Hi = Hi; p.Wi = Wi; p.M = B * Hi * Wi; p.N = N; p.K = 4 * p.Ct;
    dim3 grid(adp_cdiv(p.M, BM), adp_cdiv(p.N, BN), 4);
    igemm_kernel<<<grid, 256, 0, s>>>(p);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int simt_wgrad(int dtype, const void* s0, int M0, const void* s1, int M1, const void* g, int N, float* dw,
               int B, int Hs, int Ws, cudaStream_t s) {
  ADP_CHECK_ARG(M0 % 4 == 0 && M1 % 4 == 0 && N % 4 == 0, "wgrad: M0, M1, N must be multiples of 4");
  ADP_DISPATCH_T(dtype, {
    Wgrad<T> p;
    p.s0 = (const T*)s0; p.s1 = (const T*)s1; p.g = (const T*)g; p.dw = dw; p.M0 = M0; p.M1 = M1;
    p.M = M0 + M1; p.N = N; p.B = B; p.Hs = Hs; p.Ws = Ws; p.K = B * Hs * Ws;
    long long tiles = (long long)adp_cdiv(p.M, BM) * adp_cdiv(p.N, BN) * 16;
    int split = pick_split(tiles, p.K);
    p.chunk = adp_cdiv(adp_cdiv(p.K, split), BK) * BK;
    split = adp_cdiv(p.K, p.chunk);
    dim3 grid(adp_cdiv(p.M, BM), adp_cdiv(p.N, BN), 16 * split);
    igemm_kernel<<<grid, 256, 0, s>>>(p);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int first_conv_wgrad(int dtype, const float* x, const void* dy, float* dw, int B, int H, int W, int Cin, int N,
                     cudaStream_t s) {
  ADP_CHECK_ARG(N % 4 == 0 && H % 2 == 0 && W % 2 == 0, "first_conv_wgrad: bad shape");
  if (thin_first_supported(Cin, N)) return thin_first_conv_wgrad(dtype, x, dy, dw, B, H, W, Cin, N, s);
  ADP_DISPATCH_T(dtype, {
    FirstWgrad<T> p;
    p.x = x; p.dy = (const T*)dy; p.dw = dw; p.B = B; p.Hi = H; p.Wi = W; p.C = Cin; p.Ho = H / 2; p.Wo = W / 2;
    p.M = N; p.N = 16 * Cin; p.K = B * p.Ho * p.Wo;
    long long tiles = (long long)adp_cdiv(p.M, BM) * adp_cdiv(p.N, BN);
    int split = pick_split(tiles, p.K);
    p.chunk = adp_cdiv(adp_cdiv(p.K, split), BK) * BK;
    split = adp_cdiv(p.K, p.chunk);
    dim3 grid(adp_cdiv(p.M, BM), adp_cdiv(p.N, BN), split);
    igemm_kernel<<<grid, 256, 0, s>>>(p);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int last_convT_fprop(int dtype, const void* x0, int C0, const void* x1, int C1, const float* w,
                     const float* bias, int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s) {
  ADP_CHECK_ARG(C0 % 4 == 0 && C1 % 4 == 0 && C0 + C1 <= 2048, "last_convT: bad channel counts");
  size_t smem = (size_t)16 * (C0 + C1) * 4;
  long long npix = (long long)B * Hi * Wi;
  ADP_DISPATCH_T(dtype, {
    if (smem > 48 * 1024)
      ADP_CUDA(cudaFuncSetAttribute(last_convT_fprop_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    last_convT_fprop_kernel<T><<<pix_grid(npix), LAST_THREADS, smem, s>>>(
        (const T*)x0, C0, (const T*)x1, C1, w, bias, final_sigmoid, y, B, Hi, Wi);
  })
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int last_convT_dgrad(int dtype, const float* du, const float* w, void* g0, int C0, void* g1, int C1,
                     int B, int Hi, int Wi, cudaStream_t s) {
  ADP_CHECK_ARG(C0 % 4 == 0 && C1 % 4 == 0 && C0 + C1 <= 2048, "last_convT_dgrad: bad channel counts");
  return thin_last_convT_dgrad(dtype, du, w, g0, C0, g1, C1, B, Hi, Wi, s);
}

int last_convT_wgrad(int dtype, const void* x0, int C0, const void* x1, int C1, const float* du, float* dw,
                     int B, int Hi, int Wi, cudaStream_t s) {
  ADP_CHECK_ARG(C0 % 4 == 0 && C1 % 4 == 0, "last_convT_wgrad: bad channel counts");
  return thin_last_convT_wgrad(dtype, x0, C0, x1, C1, du, dw, B, Hi, Wi, s);
}

}  // namespace adp
