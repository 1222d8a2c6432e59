// The thin first / last layers of the U-Net on tensor cores WITHOUT a patch matrix in HBM
// (models/unetbaseline_model.py:187 outermost Conv2d(input_nc -> 64, k4 s2 p1) and :196-198 outermost
// ConvTranspose2d(128 -> 1, k4 s2 p1)).  These layers are HBM-bound: E1 reads 0.5 MB and writes 2 x 2 MB of bf16
// activations per sample, D1's backward reads / writes 4 MB.  The earlier route materialised im2col rows
// ([pixels][64] bf16, 134 MB at B = 64) with one kernel and read them back with a pointwise GEMM; here the
// [128 pixels x 64] bf16 A tile is BUILT IN SHARED MEMORY:
//   warp 0      TMA: the (2 RH + 2) x (2 CW + 8) fp32 window of the image planes that a tile of RH x CW output pixels
//               reads (zero-filled outside the image = the convolution padding), 4-deep ring; the weight tile once
//   warps 2-3   patch builders: a thread gathers the 4 x 4 x CIN windows of pixels p, p + 64 from the staged planes, converts to
//               bf16 (hi / lo split for the network input, as before) and writes row p of the SWIZZLE_128B K-major tile,
//               fence.proxy.async, mbarrier arrive
//   warp 1      tcgen05.mma (M = 128, N = 64 / 128, K = 16 per step), accumulators double-buffered in tensor memory
//   warps 4-11  two epilogue groups on alternate tiles: tcgen05.ld -> activation(s) -> swizzled staging -> full 128-byte
//               row stores (E1: LeakyReLU (centred, bordered tensor) and ReLU of the same accumulator; D1 dgrad: the two
//               64-channel halves of the concatenated input gradient)
// The weight-gradient kernel (below) accumulates  dw[m][t] += sum_pixels S[pixel][m] * patch[pixel][t]  with the same
// builder feeding the MN-major B operand, S arriving by TMA.
#include <string.h>
#include <stdlib.h>
#include "adp_tc.cuh"

namespace adp {
namespace {

using namespace tc;

constexpr int TP_THREADS = 384;             // warp 0 TMA, warp 1 MMA, warps 2-3 patch builders, warps 4-11 epilogue (2 groups)
constexpr int TP_AS = 4;                    // A-tile ring
constexpr int TP_XS = 4;                    // image-window ring
constexpr int TP_XSTAGE = 8192;             // bytes reserved per image window
constexpr int TP_A_BYTES = 128 * 128;       // [128 pixels][64 bf16]

struct ThinFwdParams {
  CUtensorMap tmImg, tmW;
  int B, Ho, Wo, CW, RH, BW, BH;            // tile = RH x CW output pixels (RH * CW = 128); window = BH x BW floats per plane
  int tiles_x, tiles_y, ntiles;
  bf16* y0; bf16* y1;
  int act;                                  // 1: y0 = lrelu(D, slope0) - center, y1 = lrelu(D, slope1) (both [pix][64]);
  float slope0, slope1;                     // 0: y0 = D[:, 0:64], y1 = D[:, 64:128]
  const float* center;
  int pad_out;                              // y0 is the interior of a [B, Ho+2, Wo+2, 64] tensor
  uint32_t box_bytes;
};

template <int NOUT>
struct ThinFwdSmem {
  static constexpr int W_BYTES = NOUT * 128;
  static constexpr int STAGING = 8 * 4096;
  static constexpr int OFF_W = TP_AS * TP_A_BYTES;
  static constexpr int OFF_X = OFF_W + W_BYTES;
  static constexpr int OFF_STG = OFF_X + TP_XS * TP_XSTAGE;
  static constexpr int OFF_BAR = OFF_STG + STAGING;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};

// Row p of the K-major SWIZZLE_128B patch tile: 16-byte chunk j of row r lives at r * 128 + ((j ^ (r & 7)) << 4).
//   CIN = 2, SPLIT: chunk kh = bf16 of the 8 values (kw, ci) of kernel row kh, chunk 4 + kh = bf16 of the rounding residue
//                   (k = (kh*4+kw)*2 + ci, and 32 + k for the residue: the layout of thin_pad_rows(.., dup = 1))
//   CIN = 1:        chunk q = kernel rows 2q, 2q+1 (k = kh*4+kw < 16); the other chunks are never read (one K step)
// xs: staged planes [CIN][BH][BW] fp32, origin = image pixel (2 ox0 - 4, 2 oy0 - 1) -- the innermost TMA start coordinate
// must be a multiple of 16 bytes (an odd float column raises "illegal instruction"); (li, lj) = pixel inside the tile
template <int CIN, bool SPLIT, bool ZERO_REST = false>
__device__ __forceinline__ void build_patch_row(const float* __restrict__ xs, int BW, int BH, int li, int lj,
                                                unsigned char* __restrict__ tile, int row) {
  static_assert((CIN == 2 && SPLIT) || (CIN == 1 && !SPLIT), "patch builder: supported input layouts");
  unsigned char* dst = tile + row * 128;
  const int sw = row & 7;
  if (CIN == 2) {
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const float* r0 = xs + (size_t)(2 * li + kh) * BW + 2 * lj + 2;     // taps kw = 0..3 are columns 2 lj + 3 .. 2 lj + 6
      const float* r1 = r0 + (size_t)BH * BW;
      const float2 a0 = *reinterpret_cast<const float2*>(r0), a1 = *reinterpret_cast<const float2*>(r0 + 2);
      const float2 a2 = *reinterpret_cast<const float2*>(r0 + 4);
      const float2 b0 = *reinterpret_cast<const float2*>(r1), b1 = *reinterpret_cast<const float2*>(r1 + 2);
      const float2 b2 = *reinterpret_cast<const float2*>(r1 + 4);
      const float w[8] = {a0.y, b0.y, a1.x, b1.x, a1.y, b1.y, a2.x, b2.x};
      float lo[8];
      uint4 h, l;
      uint32_t* hp = &h.x;
      uint32_t* lp = &l.x;
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        hp[e / 2] = pack_bf16x2(w[e], w[e + 1]);
        lo[e] = w[e] - __uint_as_float(hp[e / 2] << 16);
        lo[e + 1] = w[e + 1] - __uint_as_float(hp[e / 2] & 0xffff0000u);
        lp[e / 2] = pack_bf16x2(lo[e], lo[e + 1]);
      }
      *reinterpret_cast<uint4*>(dst + ((kh ^ sw) << 4)) = h;
      *reinterpret_cast<uint4*>(dst + (((4 + kh) ^ sw) << 4)) = l;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float* r0 = xs + (size_t)(2 * li + 2 * q) * BW + 2 * lj + 2;
      const float* r1 = r0 + BW;
      const float2 a0 = *reinterpret_cast<const float2*>(r0), a1 = *reinterpret_cast<const float2*>(r0 + 2);
      const float2 a2 = *reinterpret_cast<const float2*>(r0 + 4);
      const float2 b0 = *reinterpret_cast<const float2*>(r1), b1 = *reinterpret_cast<const float2*>(r1 + 2);
      const float2 b2 = *reinterpret_cast<const float2*>(r1 + 4);
      uint4 h;
      h.x = pack_bf16x2(a0.y, a1.x); h.y = pack_bf16x2(a1.y, a2.x);
      h.z = pack_bf16x2(b0.y, b1.x); h.w = pack_bf16x2(b1.y, b2.x);
      *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = h;
    }
    if (ZERO_REST) {      // (an operand that is read 64 columns wide: the weight-gradient GEMM)
#pragma unroll
      for (int q = 2; q < 8; ++q) *reinterpret_cast<uint4*>(dst + ((q ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// 64 accumulator columns of this lane's pixel -> bf16 -> the warp's swizzled staging tile -> 8 lanes write one 128-byte row
__device__ __forceinline__ void emit64(const float (&v)[64], unsigned char* stg, int lane, bf16* base, const unsigned (&orow)[8],
                                       bool act, float slope, const float* cen) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = act ? lrelu(v[8 * j + k], slope) : v[8 * j + k];
    if (cen) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cen + 8 * j));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(cen + 8 * j + 4));
      w[0] -= c0.x; w[1] -= c0.y; w[2] -= c0.z; w[3] -= c0.w; w[4] -= c1.x; w[5] -= c1.y; w[6] -= c1.z; w[7] -= c1.w;
    }
    uint4 u;
    u.x = pack_bf16x2(w[0], w[1]); u.y = pack_bf16x2(w[2], w[3]);
    u.z = pack_bf16x2(w[4], w[5]); u.w = pack_bf16x2(w[6], w[7]);
    *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = u;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + (lane >> 3);
    const uint4 u = *reinterpret_cast<const uint4*>(stg + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
    *reinterpret_cast<uint4*>(base + (size_t)orow[i] * 64 + (lane & 7) * 8) = u;
  }
  __syncwarp();
}

template <int CIN, bool SPLIT, int NOUT>
__global__ void __launch_bounds__(TP_THREADS, 1) thin_patch_gemm_kernel(const __grid_constant__ ThinFwdParams p) {
  using S = ThinFwdSmem<NOUT>;
  constexpr int KSTEPS = SPLIT ? 4 : 1;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* x_full = bars;                  // [XS]
  uint64_t* x_empty = x_full + TP_XS;       // [XS]
  uint64_t* a_full = x_empty + TP_XS;       // [AS]
  uint64_t* a_empty = a_full + TP_AS;       // [AS]
  uint64_t* tfull = a_empty + TP_AS;        // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint64_t* w_bar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = (int)blockIdx.x, nworkers = (int)gridDim.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmImg);
    prefetch_tmap(&p.tmW);
    for (int s = 0; s < TP_XS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 2); }
    for (int s = 0; s < TP_AS; ++s) { mbar_init(&a_full[s], 2); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * NOUT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_bar, S::W_BYTES);
      tma_load_2d(smem + S::OFF_W, &p.tmW, w_bar, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = worker; t < p.ntiles; t += nworkers) {
        const int tx = t % p.tiles_x, r = t / p.tiles_x, ty = r % p.tiles_y, b = r / p.tiles_y;
        mbar_wait(&x_empty[s], ph ^ 1);
        mbar_expect_tx(&x_full[s], p.box_bytes);
        tma_load_4d(smem + S::OFF_X + s * TP_XSTAGE, &p.tmImg, &x_full[s], 2 * tx * p.CW - 4, 2 * ty * p.RH - 1, 0, b);
        if (++s == TP_XS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, NOUT, 0, 0);
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t a_desc0 = umma_smem_desc(smem_base, 16, 1024);
      const uint64_t b_desc0 = umma_smem_desc(smem_base + S::OFF_W, 16, 1024);
      mbar_wait(w_bar, 0);
      int s = 0;
      uint32_t ph = 0, local = 0;
      for (int t = worker; t < p.ntiles; t += nworkers, ++local) {
        const uint32_t buf = local & 1u, use = local >> 1;
        mbar_wait(&tempty[buf], (use & 1u) ^ 1u);
        mbar_wait(&a_full[s], ph);
        tc_fence_after();
        const uint64_t ad = a_desc0 + (uint64_t)((uint32_t)(s * TP_A_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k)
          umma_bf16(tmem_base + buf * NOUT, ad + (uint64_t)(k * 2), b_desc0 + (uint64_t)(k * 2), idesc, k != 0 ? 1u : 0u);
        umma_commit(&a_empty[s]);
        umma_commit(&tfull[buf]);
        if (++s == TP_AS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // ===================== patch builders (two rows per thread) =====================
    const int pidx = (warp - 2) * 32 + lane;
    const int li = pidx / p.CW, lj = pidx - li * p.CW;
    const int li2 = (pidx + 64) / p.CW, lj2 = pidx + 64 - li2 * p.CW;
    int sx = 0, sa = 0;
    uint32_t phx = 0, pha = 0;
    for (int t = worker; t < p.ntiles; t += nworkers) {
      mbar_wait(&x_full[sx], phx);
      mbar_wait(&a_empty[sa], pha ^ 1);
      const float* xs = reinterpret_cast<const float*>(smem + S::OFF_X + sx * TP_XSTAGE);
      build_patch_row<CIN, SPLIT>(xs, p.BW, p.BH, li, lj, smem + sa * TP_A_BYTES, pidx);
      build_patch_row<CIN, SPLIT>(xs, p.BW, p.BH, li2, lj2, smem + sa * TP_A_BYTES, pidx + 64);
      fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a_full[sa]); mbar_arrive(&x_empty[sx]); }
      if (++sx == TP_XS) { sx = 0; phx ^= 1u; }
      if (++sa == TP_AS) { sa = 0; pha ^= 1u; }
    }
  } else {
    // ===================== epilogue: group g takes the tiles with (local & 1) == g =====================
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;
    const int pidx = q * 32 + lane;
    const int li = pidx / p.CW, lj = pidx - li * p.CW;
    unsigned char* stg = smem + S::OFF_STG + (warp - 4) * 4096;
    uint32_t local = 0;
    for (int t = worker; t < p.ntiles; t += nworkers, ++local) {
      if ((int)(local & 1u) != g) continue;
      const uint32_t use = local >> 1;
      const int tx = t % p.tiles_x, r = t / p.tiles_x, ty = r % p.tiles_y, b = r / p.tiles_y;
      const int oy = ty * p.RH + li, ox = tx * p.CW + lj;
      const unsigned opix = (unsigned)((b * p.Ho + oy) * p.Wo + ox);
      const unsigned ppix = p.pad_out ? (unsigned)((b * (p.Ho + 2) + oy + 1) * (p.Wo + 2) + ox + 1) : opix;
      unsigned orow[8], prow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        orow[i] = __shfl_sync(0xffffffffu, opix, 4 * i + (lane >> 3));
        prow[i] = __shfl_sync(0xffffffffu, ppix, 4 * i + (lane >> 3));
      }
      mbar_wait(&tfull[g], use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + g * NOUT + ((uint32_t)(q * 32) << 16);
      float v[64];
      tmem_ld32(tacc, v);
      tmem_ld32(tacc + 32u, v + 32);
      if (p.act) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[g]);           // (the accumulator is in registers: the next MMA may overwrite it)
        emit64(v, stg, lane, p.y0, prow, true, p.slope0, p.center);
        emit64(v, stg, lane, p.y1, orow, true, p.slope1, nullptr);
      } else {
        emit64(v, stg, lane, p.y0, orow, false, 0.f, nullptr);
        if (NOUT > 64) {
          tmem_ld32(tacc + 64u, v);
          tmem_ld32(tacc + 96u, v + 32);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[g]);
        if (NOUT > 64) emit64(v, stg, lane, p.y1, orow, false, 0.f, nullptr);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * NOUT);
  }
}

// ------------------------------------------------------------------ D1 forward: ConvTranspose2d(128 -> 1) + bias + ReLU | Sigmoid
// y[b, 2i+a, 2j+c] = act(bias + sum over the 2 x 2 taps that land there of P[b, i', j'][kh*4+kw]),  P[pix][t] = sum_ch (r|q)[pix][ch] w[ch][t].
// One input row (Wi = 128 pixels) is one M = 128 tile: P_i = [128 x 128 ch] . [128 ch x 16 taps] on the tensor core (N = 16,
// 8 K steps).  A CTA walks a band of consecutive rows; the epilogue keeps the last three P rows in shared memory, so output
// rows 2i-1 and 2i are finished as soon as P_i exists (plus 2i+1 at the bottom of an image) -- P never goes to HBM and
// there is no col2im pass.  The second input half may arrive as t, the raw output of the previous transposed conv: the
// BatchNorm + ReLU that makes q = ReLU(t * scale + shift) is applied to the tile in shared memory (then q is never written
// by the forward pass either).
constexpr int TL_THREADS = 320;             // warp 0 TMA, warp 1 MMA, warps 2-5 tile transform, warps 6-9 epilogue
constexpr int TL_STAGES = 4;
constexpr int TL_HALF = 128 * 128;          // [128 pixels][64 ch] bf16
constexpr int TL_PSTRIDE = 20;              // floats per pixel in the P ring (16 taps + padding: conflict-free float4 rows)

struct ThinLastFwdParams {
  CUtensorMap tmR, tmT, tmW;
  int B, Hi, Wi, rows;
  const float* bn_scale; const float* bn_shift;     // NULL: the second tensor already holds q
  const float* bias; int final_sigmoid;
  float* y;
};

struct ThinLastSmem {
  static constexpr int OFF_W = TL_STAGES * 2 * TL_HALF;          // two [16 taps][64 ch] tiles
  static constexpr int OFF_P = OFF_W + 2 * 2048;                 // 3 x [128 pixels][TL_PSTRIDE] floats
  static constexpr int OFF_TAB = OFF_P + 3 * 128 * TL_PSTRIDE * 4;
  static constexpr int OFF_BAR = OFF_TAB + 2 * 64 * 4;
  static constexpr int BYTES = OFF_BAR + 256 + 1024;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(TL_THREADS, 1) thin_last_fwd_kernel(const __grid_constant__ ThinLastFwdParams p) {
  using S = ThinLastSmem;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* full = bars;                     // [STAGES] TMA bytes
  uint64_t* ready = full + TL_STAGES;        // [STAGES] tile transformed (4 warps)
  uint64_t* empty = ready + TL_STAGES;       // [STAGES]
  uint64_t* tfull = empty + TL_STAGES;       // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint64_t* w_bar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* tab = reinterpret_cast<float*>(smem + S::OFF_TAB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's band of rows g = b * Hi + i, preceded by one halo row (P only) when the band starts inside an image
  const int g_begin = (int)((long long)blockIdx.x * p.rows / gridDim.x), g_end = (int)((long long)(blockIdx.x + 1) * p.rows / gridDim.x);
  const int halo = (g_begin % p.Hi) != 0 ? 1 : 0;
  const int g_first = g_begin - halo, ntiles = g_end - g_first;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmR);
    prefetch_tmap(&p.tmT);
    prefetch_tmap(&p.tmW);
    for (int s = 0; s < TL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  if (p.bn_scale && threadIdx.x >= 64 && threadIdx.x < 128) {
    tab[threadIdx.x - 64] = p.bn_scale[threadIdx.x - 64];
    tab[threadIdx.x] = p.bn_shift[threadIdx.x - 64];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_bar, 2 * 2048);
      tma_load_2d(smem + S::OFF_W, &p.tmW, w_bar, 0, 0);
      tma_load_2d(smem + S::OFF_W + 2048, &p.tmW, w_bar, 64, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int k = 0; k < ntiles; ++k) {
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* dst = smem + s * 2 * TL_HALF;
        mbar_expect_tx(&full[s], 2 * TL_HALF);
        tma_load_3d(dst, &p.tmR, &full[s], 0, 0, g_first + k);
        tma_load_3d(dst + TL_HALF, &p.tmT, &full[s], 0, 0, g_first + k);
        if (++s == TL_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t b_desc0 = umma_smem_desc(smem_base + S::OFF_W, 16, 1024);
      mbar_wait(w_bar, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int k = 0; k < ntiles; ++k) {
        const uint32_t buf = (uint32_t)k & 1u, use = (uint32_t)k >> 1;
        mbar_wait(&tempty[buf], (use & 1u) ^ 1u);
        mbar_wait(&ready[s], ph);
        tc_fence_after();
        const uint64_t a_desc0 = umma_smem_desc(smem_base + s * 2 * TL_HALF, 16, 1024);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + buf * 16, a_desc0 + (uint64_t)((h * TL_HALF) >> 4) + (uint64_t)(kk * 2),
                      b_desc0 + (uint64_t)((h * 2048) >> 4) + (uint64_t)(kk * 2), idesc, (h | kk) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
        umma_commit(&tfull[buf]);
        if (++s == TL_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ===================== q = ReLU(t * scale + shift) on the staged tile, in place =====================
    const int tid = (warp - 2) * 32 + lane;
    // thread tid handles the 16-byte chunks c = it * 128 + tid: row c >> 3 = it * 16 + (tid >> 3), physical position tid & 7, so
    // its 8 channels (SWIZZLE_128B: logical chunk = physical ^ (row & 7)) are the same for every chunk: coefficients in registers
    const int ch0 = (((tid & 7) ^ ((tid >> 3) & 7)) << 3);
    float4 sc0 = make_float4(0.f, 0.f, 0.f, 0.f), sc1 = sc0, sh0 = sc0, sh1 = sc0;
    if (p.bn_scale) {
      sc0 = *reinterpret_cast<const float4*>(tab + ch0); sc1 = *reinterpret_cast<const float4*>(tab + ch0 + 4);
      sh0 = *reinterpret_cast<const float4*>(tab + 64 + ch0); sh1 = *reinterpret_cast<const float4*>(tab + 64 + ch0 + 4);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int k = 0; k < ntiles; ++k) {
      mbar_wait(&full[s], ph);
      if (p.bn_scale) {
        uint4* tq = reinterpret_cast<uint4*>(smem + s * 2 * TL_HALF + TL_HALF);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int c = it * 128 + tid;
          const float8 t = cvt8(tq[c]);
          uint4 u;
          u.x = pack_bf16x2(fmaxf(fmaf(t.v[0], sc0.x, sh0.x), 0.f), fmaxf(fmaf(t.v[1], sc0.y, sh0.y), 0.f));
          u.y = pack_bf16x2(fmaxf(fmaf(t.v[2], sc0.z, sh0.z), 0.f), fmaxf(fmaf(t.v[3], sc0.w, sh0.w), 0.f));
          u.z = pack_bf16x2(fmaxf(fmaf(t.v[4], sc1.x, sh1.x), 0.f), fmaxf(fmaf(t.v[5], sc1.y, sh1.y), 0.f));
          u.w = pack_bf16x2(fmaxf(fmaf(t.v[6], sc1.z, sh1.z), 0.f), fmaxf(fmaf(t.v[7], sc1.w, sh1.w), 0.f));
          tq[c] = u;
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[s]);
      if (++s == TL_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: lane = pixel j of the row =====================
    const int q = warp & 3;
    const int j = q * 32 + lane;
    float* pring = reinterpret_cast<float*>(smem + S::OFF_P);
    const float bias = p.bias ? p.bias[0] : 0.f;
    const int Wo = 2 * p.Wi, Ho = 2 * p.Hi;
    for (int k = 0; k < ntiles; ++k) {
      const uint32_t buf = (uint32_t)k & 1u, use = (uint32_t)k >> 1;
      const int g = g_first + k, b = g / p.Hi, i = g - b * p.Hi;
      mbar_wait(&tfull[buf], use & 1u);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem_base + buf * 16 + ((uint32_t)(q * 32) << 16), v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      float* cur = pring + (k % 3) * 128 * TL_PSTRIDE;
      const float* prev = pring + ((k + 2) % 3) * 128 * TL_PSTRIDE;
#pragma unroll
      for (int t4 = 0; t4 < 16; t4 += 4)
        *reinterpret_cast<float4*>(cur + j * TL_PSTRIDE + t4) = make_float4(v[t4], v[t4 + 1], v[t4 + 2], v[t4 + 3]);
      asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps: P_i complete (3 slots: one barrier per row)
      if (halo && k == 0) continue;                          // (P only: the rows this one finishes belong to the previous band)
      // output (oy, 2j + c): rows (i', kh) and columns (j', kw) of the taps that land there
      auto emit_row = [&](int oy, const float* pa, int kha, const float* pb, int khb) {
        float u0 = bias, u1 = bias;                          // c = 0: (j-1, kw 3), (j, kw 1);  c = 1: (j, kw 2), (j+1, kw 0)
        if (pa) {
          if (j > 0) u0 += pa[(j - 1) * TL_PSTRIDE + kha * 4 + 3];
          u0 += pa[j * TL_PSTRIDE + kha * 4 + 1];
          u1 += pa[j * TL_PSTRIDE + kha * 4 + 2];
          if (j + 1 < p.Wi) u1 += pa[(j + 1) * TL_PSTRIDE + kha * 4 + 0];
        }
        if (pb) {
          if (j > 0) u0 += pb[(j - 1) * TL_PSTRIDE + khb * 4 + 3];
          u0 += pb[j * TL_PSTRIDE + khb * 4 + 1];
          u1 += pb[j * TL_PSTRIDE + khb * 4 + 2];
          if (j + 1 < p.Wi) u1 += pb[(j + 1) * TL_PSTRIDE + khb * 4 + 0];
        }
        if (p.final_sigmoid) { u0 = 1.f / (1.f + expf(-u0)); u1 = 1.f / (1.f + expf(-u1)); }
        else { u0 = fmaxf(u0, 0.f); u1 = fmaxf(u1, 0.f); }
        *reinterpret_cast<float2*>(p.y + ((size_t)b * Ho + oy) * Wo + 2 * j) = make_float2(u0, u1);
      };
      // oy = 2i - 1 (a = 1 of row i-1): rows i-1 (kh 2) and i (kh 0);  oy = 2i (a = 0): rows i-1 (kh 3) and i (kh 1)
      if (i > 0) emit_row(2 * i - 1, prev, 2, cur, 0);
      emit_row(2 * i, i > 0 ? prev : nullptr, 3, cur, 1);
      if (i == p.Hi - 1) emit_row(2 * i + 1, cur, 2, nullptr, 0);      // bottom edge: row i (kh 2) only
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// fp32 image planes [B][CIN][H][W] as a 4-D tensor map (W | H | CIN | B), no swizzle, zero fill outside
int make_tmap_image(CUtensorMap* out, const float* img, int B, int CIN, int H, int W, int BW, int BH) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { adp_set_error("cuTensorMapEncodeTiled is not available from the driver"); return ADP_ERR_CUDA; }
  cuuint64_t gd[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CIN, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)CIN * H * W * 4};
  cuuint32_t bx[4] = {(cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)CIN, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(img), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    adp_set_error("cuTensorMapEncodeTiled (image planes) failed (%d): [%d,%d,%d,%d] box [%d,%d]", (int)r, W, H, CIN, B, BW, BH);
    return ADP_ERR_CUDA;
  }
  return ADP_OK;
}

bool pow2i(int v) { return v > 0 && (v & (v - 1)) == 0; }

// tile geometry over the Ho x Wo output grid of a k4 s2 p1 window walk
bool thin_geometry(int Ho, int Wo, int CIN, int* CW, int* RH, int* BW, int* BH) {
  if (!pow2i(Ho) || !pow2i(Wo) || Wo < 16) return false;
  *CW = Wo < 64 ? Wo : 64;
  *RH = 128 / *CW;
  if (*RH > Ho) return false;
  *BW = 2 * *CW + 8;
  *BH = 2 * *RH + 2;
  return (size_t)*BW * *BH * CIN * 4 <= (size_t)TP_XSTAGE;
}

template <int CIN, bool SPLIT, int NOUT>
int launch_thin_fwd(ThinFwdParams& p, cudaStream_t s) {
  using S = ThinFwdSmem<NOUT>;
  ADP_SMEM_ATTR((thin_patch_gemm_kernel<CIN, SPLIT, NOUT>), S::BYTES);
  const int ctas = p.ntiles < sm_count() ? p.ntiles : sm_count();
  thin_patch_gemm_kernel<CIN, SPLIT, NOUT><<<ctas, TP_THREADS, S::BYTES, s>>>(p);
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int fill_fwd_params(ThinFwdParams* p, const float* img, int B, int CIN, int H, int W, const void* w_nk, int NOUT) {
  memset(p, 0, sizeof(*p));
  const int Ho = H / 2, Wo = W / 2;
  ADP_CHECK_ARG(thin_geometry(Ho, Wo, CIN, &p->CW, &p->RH, &p->BW, &p->BH), "thin tc: unsupported image size %dx%d", H, W);
  ADP_CHECK_ARG((long long)B * (Ho + 2) * (Wo + 2) < (1LL << 31), "thin tc: too many pixels");
  p->B = B; p->Ho = Ho; p->Wo = Wo;
  p->tiles_x = Wo / p->CW; p->tiles_y = Ho / p->RH; p->ntiles = B * p->tiles_x * p->tiles_y;
  p->box_bytes = (uint32_t)(p->BW * p->BH * CIN * 4);
  ADP_TRY(make_tmap_image(&p->tmImg, img, B, CIN, H, W, p->BW, p->BH));
  uint64_t dims[2] = {64, (uint64_t)NOUT};
  uint64_t str[1] = {64 * 2};
  uint32_t box[2] = {64, (uint32_t)NOUT};
  return make_tmap_bf16(&p->tmW, w_nk, 2, dims, str, box);
}


// ------------------------------------------------------------------ weight gradients
//   FOLD = false (D1):  dw[c][t] += sum_pix (s0|s1)[pix][c] * patch[pix][t]          c < 128 (two 64-channel tensors), t < 16
//   FOLD = true  (E1):  dw[n][t] += sum_pix s0[pix][n] * (patch_hi + patch_lo)[pix][t]   n < 64, t < 32: pixel PAIRS are the
//                       GEMM's K rows, D[(h,n)][(h',t)] with the two diagonal blocks h = h' summed (M = 128 without padding)
// Both operands are MN-major (rows of the shared-memory tiles = pixels = K).  One accumulator per CTA over all its tiles,
// then red.global.add into dw (zeroed by the caller).
constexpr int TW_THREADS = 192;             // warp 0 TMA, warp 1 MMA, warps 2-5 patch builders + final epilogue
constexpr int TW_STAGES = 3;

struct ThinWgradParams {
  CUtensorMap tmImg, tmS0, tmS1, tmS2;
  float slope;                              // XF 1: S = gA * (r > 0 ? 1 : slope) + (r > 0 ? gB : 0), formed in shared memory
  const float* bn_scale; const float* bn_shift;   // XF 2: the second S half arrives as t, q = ReLU(t * scale + shift) is formed in shared memory
  int B, Ho, Wo, CW, RH, BW, BH;
  int tiles_x, tiles_y, ntiles;
  float* dw;
  uint32_t box_bytes;
};

template <bool FOLD, int XF = 0>
struct ThinWgradSmem {
  static constexpr bool ACT = XF == 1;
  static constexpr int KROWS = FOLD ? 64 : 128;
  static constexpr int HALF = KROWS * 128;                 // one [KROWS][64] bf16 region
  static constexpr int A_BYTES = 2 * HALF;
  static constexpr int B_BYTES = (FOLD ? 2 : 1) * HALF;
  static constexpr int EXTRA = ACT ? 2 * A_BYTES : 0;      // ACT: the tiles of gB and r next to gA's (which becomes S in place)
  static constexpr int STAGE = A_BYTES + B_BYTES + EXTRA;
  static constexpr int STAGES = ACT ? 2 : TW_STAGES;
  static constexpr int OFF_X = STAGES * STAGE;
  static constexpr int OFF_BAR = OFF_X + TP_XS * TP_XSTAGE;
  static constexpr int OFF_TAB = OFF_BAR + 256;            // XF 2: scale | shift, 2 x 64 floats
  static constexpr int BYTES = OFF_TAB + 512 + 1024;
  static constexpr int NT = FOLD ? 128 : 64;
};

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool FOLD, int XF = 0>
__global__ void __launch_bounds__(TW_THREADS, 1) thin_patch_wgrad_kernel(const __grid_constant__ ThinWgradParams p) {
  constexpr bool ACT = XF == 1, QBN = XF == 2;
  static_assert(!ACT || FOLD, "the fused activation backward belongs to the first conv's weight gradient");
  static_assert(!QBN || !FOLD, "the BatchNorm + ReLU on load belongs to the last transposed conv's weight gradient");
  using S = ThinWgradSmem<FOLD, XF>;
  constexpr int STAGES = S::STAGES;
  constexpr int KSTEPS = S::KROWS / 16;
  constexpr int NT = S::NT;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* x_full = bars;                  // [XS]
  uint64_t* x_empty = x_full + TP_XS;       // [XS]
  uint64_t* full = x_empty + TP_XS;         // [STAGES]  TMA bytes of S + the 4 builder warps (ACT: the builder warps only)
  uint64_t* empty = full + TW_STAGES;       // [STAGES]
  uint64_t* loaded = empty + TW_STAGES;     // [STAGES]  ACT: the three source tiles have landed
  uint64_t* accum_bar = loaded + TW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = (int)blockIdx.x, nworkers = (int)gridDim.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmImg);
    prefetch_tmap(&p.tmS0);
    if (!FOLD || ACT) prefetch_tmap(&p.tmS1);
    if (ACT) prefetch_tmap(&p.tmS2);
    for (int s = 0; s < TP_XS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 4); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], (ACT || QBN) ? 4 : 5); mbar_init(&empty[s], 1); mbar_init(&loaded[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NT);
  float* tab = reinterpret_cast<float*>(smem + S::OFF_TAB);
  if (QBN && threadIdx.x >= 64 && threadIdx.x < 128) {
    tab[threadIdx.x - 64] = p.bn_scale[threadIdx.x - 64];
    tab[threadIdx.x] = p.bn_shift[threadIdx.x - 64];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int sx = 0, s = 0;
      uint32_t phx = 0, ph = 0;
      for (int t = worker; t < p.ntiles; t += nworkers) {
        const int tx = t % p.tiles_x, r = t / p.tiles_x, ty = r % p.tiles_y, b = r / p.tiles_y;
        mbar_wait(&x_empty[sx], phx ^ 1);
        mbar_expect_tx(&x_full[sx], p.box_bytes);
        tma_load_4d(smem + S::OFF_X + sx * TP_XSTAGE, &p.tmImg, &x_full[sx], 2 * tx * p.CW - 4, 2 * ty * p.RH - 1, 0, b);
        if (++sx == TP_XS) { sx = 0; phx ^= 1u; }
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* a_dst = smem + s * S::STAGE;
        if (ACT) {
          mbar_expect_tx(&loaded[s], 3 * S::A_BYTES);
          unsigned char* e_dst = a_dst + S::A_BYTES + S::B_BYTES;
          const int cx = tx * (p.CW / 2), cy = b * p.Ho + ty * p.RH;
          tma_load_3d(a_dst, &p.tmS0, &loaded[s], 0, cx, cy);
          tma_load_3d(a_dst + S::HALF, &p.tmS0, &loaded[s], 64, cx, cy);
          tma_load_3d(e_dst, &p.tmS1, &loaded[s], 0, cx, cy);
          tma_load_3d(e_dst + S::HALF, &p.tmS1, &loaded[s], 64, cx, cy);
          tma_load_3d(e_dst + S::A_BYTES, &p.tmS2, &loaded[s], 0, cx, cy);
          tma_load_3d(e_dst + S::A_BYTES + S::HALF, &p.tmS2, &loaded[s], 64, cx, cy);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          continue;
        }
        uint64_t* a_bar = QBN ? &loaded[s] : &full[s];
        mbar_expect_tx(a_bar, S::A_BYTES);
        if (FOLD) {
          tma_load_3d(a_dst, &p.tmS0, &full[s], 0, tx * (p.CW / 2), b * p.Ho + ty * p.RH);
          tma_load_3d(a_dst + S::HALF, &p.tmS0, &full[s], 64, tx * (p.CW / 2), b * p.Ho + ty * p.RH);
        } else {
          tma_load_3d(a_dst, &p.tmS0, a_bar, 0, tx * p.CW, b * p.Ho + ty * p.RH);
          tma_load_3d(a_dst + S::HALF, &p.tmS1, a_bar, 0, tx * p.CW, b * p.Ho + ty * p.RH);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, NT, 1, 1);
      const uint32_t smem_base = smem_u32(smem);
      int s = 0;
      uint32_t ph = 0, first = 1;
      for (int t = worker; t < p.ntiles; t += nworkers) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * S::STAGE, b_addr = a_addr + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          const uint64_t ad = umma_smem_desc(a_addr + k * 2048, S::HALF, 1024);
          const uint64_t bd = umma_smem_desc(b_addr + k * 2048, S::HALF, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
        }
        first = 0;
        umma_commit(&empty[s]);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(accum_bar);
    }
  } else {
    const int pidx = (warp - 2) * 32 + lane;
    const int li = pidx / p.CW, lj = pidx - li * p.CW;
    int sx = 0, s = 0;
    uint32_t phx = 0, ph = 0;
    for (int t = worker; t < p.ntiles; t += nworkers) {
      mbar_wait(&x_full[sx], phx);
      mbar_wait(&empty[s], ph ^ 1);
      unsigned char* b_dst = smem + s * S::STAGE + S::A_BYTES;
      const float* xs = reinterpret_cast<const float*>(smem + S::OFF_X + sx * TP_XSTAGE);
      if (ACT) {
        // S = dL/de of the first conv, formed in place from the three tiles (same swizzled layout: pure element-wise):
        //   gz = gA * lrelu'(e, slope) + gB * relu'(e),  e > 0 <=> r = ReLU(e) > 0   (models/unetbaseline_model.py:187-192)
        mbar_wait(&loaded[s], ph);
        uint4* ga = reinterpret_cast<uint4*>(smem + s * S::STAGE);
        const uint4* gb = reinterpret_cast<const uint4*>(smem + s * S::STAGE + S::A_BYTES + S::B_BYTES);
        const uint4* rr = gb + S::A_BYTES / 16;
#pragma unroll
        for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) {
          const int c = i * 128 + pidx;
          const float8 a = cvt8(ga[c]), bq = cvt8(gb[c]), r8 = cvt8(rr[c]);
          float8 o;
#pragma unroll
          for (int k = 0; k < 8; ++k) o.v[k] = r8.v[k] > 0.f ? a.v[k] + bq.v[k] : a.v[k] * p.slope;
          uint4 u;
          u.x = pack_bf16x2(o.v[0], o.v[1]); u.y = pack_bf16x2(o.v[2], o.v[3]);
          u.z = pack_bf16x2(o.v[4], o.v[5]); u.w = pack_bf16x2(o.v[6], o.v[7]);
          ga[c] = u;
        }
      }
      if (QBN) {
        // q = ReLU(t * scale + shift) on the second S half, in place (chunk at physical position c & 7 of row c >> 3 holds
        // the channels 8 * ((c & 7) ^ (row & 7)) ..)
        mbar_wait(&loaded[s], ph);
        uint4* tq = reinterpret_cast<uint4*>(smem + s * S::STAGE + S::HALF);
        // (row & 7 of chunk it * 128 + pidx is (pidx >> 3) & 7 for every it: one set of 8 channels per thread)
        const int ch0 = (((pidx & 7) ^ ((pidx >> 3) & 7)) << 3);
        const float4 sc0 = *reinterpret_cast<const float4*>(tab + ch0), sc1 = *reinterpret_cast<const float4*>(tab + ch0 + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(tab + 64 + ch0), sh1 = *reinterpret_cast<const float4*>(tab + 64 + ch0 + 4);
#pragma unroll
        for (int it = 0; it < S::HALF / 16 / 128; ++it) {
          const int c = it * 128 + pidx;
          const float8 t = cvt8(tq[c]);
          uint4 u;
          u.x = pack_bf16x2(fmaxf(fmaf(t.v[0], sc0.x, sh0.x), 0.f), fmaxf(fmaf(t.v[1], sc0.y, sh0.y), 0.f));
          u.y = pack_bf16x2(fmaxf(fmaf(t.v[2], sc0.z, sh0.z), 0.f), fmaxf(fmaf(t.v[3], sc0.w, sh0.w), 0.f));
          u.z = pack_bf16x2(fmaxf(fmaf(t.v[4], sc1.x, sh1.x), 0.f), fmaxf(fmaf(t.v[5], sc1.y, sh1.y), 0.f));
          u.w = pack_bf16x2(fmaxf(fmaf(t.v[6], sc1.z, sh1.z), 0.f), fmaxf(fmaf(t.v[7], sc1.w, sh1.w), 0.f));
          tq[c] = u;
        }
      }
      if (FOLD) build_patch_row<2, true>(xs, p.BW, p.BH, li, lj, b_dst + (lj & 1) * S::HALF, li * (p.CW / 2) + (lj >> 1));
      else build_patch_row<1, false, true>(xs, p.BW, p.BH, li, lj, b_dst, pidx);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&full[s]); mbar_arrive(&x_empty[sx]); }
      if (++sx == TP_XS) { sx = 0; phx ^= 1u; }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    // ---- final epilogue: accumulator row m = q * 32 + lane
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    if (worker < p.ntiles) {
      if (FOLD) {
        const int h = m >> 6, n = m & 63;
        float v[64];
        tmem_ld32(tacc + (uint32_t)(h * 64), v);
        tmem_ld32(tacc + (uint32_t)(h * 64 + 32), v + 32);
        float* dst = p.dw + n * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          red_add_f32x4(dst + i, v[i] + v[32 + i], v[i + 1] + v[33 + i], v[i + 2] + v[34 + i], v[i + 3] + v[35 + i]);
      } else {
        float v[32];
        tmem_ld32(tacc, v);
        float* dst = p.dw + m * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4) red_add_f32x4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NT);
  }
}

template <bool FOLD, int XF = 0>
int launch_thin_wgrad(ThinWgradParams& p, cudaStream_t s) {
  using S = ThinWgradSmem<FOLD, XF>;
  ADP_SMEM_ATTR((thin_patch_wgrad_kernel<FOLD, XF>), S::BYTES);
  const int ctas = p.ntiles < sm_count() ? p.ntiles : sm_count();
  thin_patch_wgrad_kernel<FOLD, XF><<<ctas, TW_THREADS, S::BYTES, s>>>(p);
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

int fill_wgrad_params(ThinWgradParams* p, const float* img, int B, int CIN, int H, int W, float* dw) {
  memset(p, 0, sizeof(*p));
  const int Ho = H / 2, Wo = W / 2;
  ADP_CHECK_ARG(thin_geometry(Ho, Wo, CIN, &p->CW, &p->RH, &p->BW, &p->BH), "thin tc wgrad: unsupported image size %dx%d", H, W);
  p->B = B; p->Ho = Ho; p->Wo = Wo; p->dw = dw;
  p->tiles_x = Wo / p->CW; p->tiles_y = Ho / p->RH; p->ntiles = B * p->tiles_x * p->tiles_y;
  p->box_bytes = (uint32_t)(p->BW * p->BH * CIN * 4);
  return make_tmap_image(&p->tmImg, img, B, CIN, H, W, p->BW, p->BH);
}

}  // namespace

bool thin_tc_supported(int B, int Cin, int H, int W) {
  int CW, RH, BW, BH;
  if (!adp_device_is_sm100() || !encode_tiled_fn() || B < 1 || H % 2 || W % 2) return false;
  return (Cin == 1 || Cin == 2) && thin_geometry(H / 2, W / 2, Cin, &CW, &RH, &BW, &BH);
}

// E1 forward: x fp32 [B,2,H,W] -> a = lrelu(conv, slope0) - center (interior of the bordered tensor when pad_out),
// r = lrelu(conv, slope1); w_pad = bf16 [64][64] from thin_pad_rows(conv_w, 64, 32, dup = 1)
int thin_tc_first_conv(const float* x, const void* w_pad, void* a, float slope0, void* r, float slope1, const float* center,
                       int pad_out, int B, int H, int W, cudaStream_t s) {
  ThinFwdParams p;
  ADP_TRY(fill_fwd_params(&p, x, B, 2, H, W, w_pad, 64));
  p.y0 = reinterpret_cast<bf16*>(a); p.y1 = reinterpret_cast<bf16*>(r);
  p.act = 1; p.slope0 = slope0; p.slope1 = slope1; p.center = center; p.pad_out = pad_out ? 1 : 0;
  return launch_thin_fwd<2, true, 64>(p, s);
}

// D1 data gradient: du fp32 [B,1,2Hi,2Wi] -> g0 | g1 = [pix][64] each; w_pad = bf16 [128][64] from thin_pad_rows(convT_w, 128, 16, 0)
int thin_tc_last_dgrad(const float* du, const void* w_pad, void* g0, void* g1, int B, int Hi, int Wi, cudaStream_t s) {
  ThinFwdParams p;
  ADP_TRY(fill_fwd_params(&p, du, B, 1, 2 * Hi, 2 * Wi, w_pad, 128));
  p.y0 = reinterpret_cast<bf16*>(g0); p.y1 = reinterpret_cast<bf16*>(g1);
  return launch_thin_fwd<1, false, 128>(p, s);
}

bool thin_tc_last_fwd_supported(int B, int Hi, int Wi) {
  return adp_device_is_sm100() && encode_tiled_fn() && B >= 1 && Hi >= 1 && Wi == 128 && (long long)B * Hi < (1LL << 30);
}

// D1 forward: x0 = r, x1 = t (scale / shift given: q = ReLU(t * scale + shift) is formed in shared memory) or q (scale NULL);
// both bf16 [B,Hi,128,64]; w16: bf16 [16 taps][128 ch] (cast_transpose_taps); y fp32 [B,1,2Hi,256] = act(bias + convT)
int thin_tc_last_fwd(const void* x0, const void* x1, const float* scale, const float* shift, const void* w16, const float* bias,
                     int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s) {
  ADP_CHECK_ARG(thin_tc_last_fwd_supported(B, Hi, Wi), "thin_tc_last_fwd: unsupported shape %dx%dx%d (needs Wi = 128)", B, Hi, Wi);
  ThinLastFwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Hi = Hi; p.Wi = Wi; p.rows = B * Hi;
  p.bn_scale = scale; p.bn_shift = scale ? shift : nullptr;
  p.bias = bias; p.final_sigmoid = final_sigmoid; p.y = y;
  for (int h = 0; h < 2; ++h) {
    uint64_t dims[3] = {64, (uint64_t)Wi, (uint64_t)B * Hi};
    uint64_t str[2] = {64 * 2, (uint64_t)Wi * 64 * 2};
    uint32_t box[3] = {64, 128, 1};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmR : &p.tmT, h == 0 ? x0 : x1, 3, dims, str, box));
  }
  {
    uint64_t dims[2] = {128, 16};
    uint64_t str[1] = {128 * 2};
    uint32_t box[2] = {64, 16};
    ADP_TRY(make_tmap_bf16(&p.tmW, w16, 2, dims, str, box));
  }
  ADP_SMEM_ATTR(thin_last_fwd_kernel, ThinLastSmem::BYTES);
  const int ctas = p.rows < sm_count() ? p.rows : sm_count();
  thin_last_fwd_kernel<<<ctas, TL_THREADS, ThinLastSmem::BYTES, s>>>(p);
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

// E1 weight gradient: dw fp32 [64][16][2] += sum_pix g_e[pix][n] * x-patch; g_e bf16 [B,H/2,W/2,64]; dw zeroed by the caller
int thin_tc_first_wgrad(const float* x, const void* g_e, float* dw, int B, int H, int W, cudaStream_t s) {
  ThinWgradParams p;
  ADP_TRY(fill_wgrad_params(&p, x, B, 2, H, W, dw));
  uint64_t dims[3] = {128, (uint64_t)p.Wo / 2, (uint64_t)B * p.Ho};
  uint64_t str[2] = {128 * 2, (uint64_t)(p.Wo / 2) * 128 * 2};
  uint32_t box[3] = {64, (uint32_t)p.CW / 2, (uint32_t)p.RH};
  ADP_TRY(make_tmap_bf16(&p.tmS0, g_e, 3, dims, str, box));
  return launch_thin_wgrad<true>(p, s);
}

// E1 weight gradient with the activation backward of level 0 folded in: dL/de = gA * lrelu'(e, slope) + gB * relu'(e) is
// formed tile by tile in shared memory from gA, gB and r = ReLU(e) (all bf16 [B,H/2,W/2,64]) and never written to HBM
int thin_tc_first_wgrad_act(const float* x, const void* gA, const void* gB, const void* r, float slope, float* dw, int B, int H,
                            int W, cudaStream_t s) {
  ThinWgradParams p;
  ADP_TRY(fill_wgrad_params(&p, x, B, 2, H, W, dw));
  p.slope = slope;
  const void* src[3] = {gA, gB, r};
  CUtensorMap* maps[3] = {&p.tmS0, &p.tmS1, &p.tmS2};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[3] = {128, (uint64_t)p.Wo / 2, (uint64_t)B * p.Ho};
    uint64_t str[2] = {128 * 2, (uint64_t)(p.Wo / 2) * 128 * 2};
    uint32_t box[3] = {64, (uint32_t)p.CW / 2, (uint32_t)p.RH};
    ADP_TRY(make_tmap_bf16(maps[i], src[i], 3, dims, str, box));
  }
  return launch_thin_wgrad<true, 1>(p, s);
}

// D1 weight gradient: dw fp32 [128][16] += sum_pix (x0|x1)[pix][c] * du-patch; x0, x1 bf16 [B,Hi,Wi,64]; dw zeroed by the caller
int thin_tc_last_wgrad(const void* x0, const void* x1, const float* du, float* dw, int B, int Hi, int Wi, cudaStream_t s,
                       const float* x1_scale, const float* x1_shift) {
  ThinWgradParams p;
  ADP_TRY(fill_wgrad_params(&p, du, B, 1, 2 * Hi, 2 * Wi, dw));
  p.bn_scale = x1_scale; p.bn_shift = x1_shift;
  for (int h = 0; h < 2; ++h) {
    uint64_t dims[3] = {64, (uint64_t)p.Wo, (uint64_t)B * p.Ho};
    uint64_t str[2] = {64 * 2, (uint64_t)p.Wo * 64 * 2};
    uint32_t box[3] = {64, (uint32_t)p.CW, (uint32_t)p.RH};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmS0 : &p.tmS1, h == 0 ? x0 : x1, 3, dims, str, box));
  }
  if (x1_scale) return launch_thin_wgrad<false, 2>(p, s);      // x1 = t: q = ReLU(t * scale + shift) formed on load
  return launch_thin_wgrad<false>(p, s);
}

}  // namespace adp
