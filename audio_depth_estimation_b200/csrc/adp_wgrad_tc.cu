// Weight gradients of the k4-s2-p1 convolutions on tcgen05 tensor cores (family F3):
//   dw[m][kh,kw][n] += sum_{b,i,j} S[b,i,j,m] * G[b,2i-1+kh,2j-1+kw,n]
// (Conv2d: S = dL/dy, G = layer input;  ConvTranspose2d: S = layer input (r|q concat), G = dL/dy.)
// Replaces cudnnConvolutionBackwardFilter behind models/unetbaseline_model.py:187,:196,:209,:218.
//
// Per filter tap this is a GEMM whose reduction index is the PIXEL, while both NHWC operands are
// contiguous in their channel index -- i.e. both operands are "MN-major".  tcgen05 reads such
// operands directly through MN-major SWIZZLE_128B shared-memory descriptors, so there is no
// transpose pass: TMA drops [64 pixels x 64 channels] boxes of S and of the tap-shifted view of G
// (same 5-D stride-2 view and zero-filled halo as the forward conv) into smem, and one CTA
// accumulates a [128 x NT] weight tile for FOUR taps at once in tensor memory (4*NT <= 512
// columns), re-using the S tile across the taps.  The pixel range is split across CTAs until
// the grid fills the GPU; partial tiles are reduced with fp32 red.global.add.
#include <stdlib.h>
#include "adp_tc.cuh"

namespace adp {
namespace {

using namespace tc;

constexpr int WG_M = 128;            // weight-tile rows (channels of S)
constexpr int WG_P = 32;             // pixels per k-block (two UMMA K steps): small stages -> a deep TMA ring
constexpr int WG_TAPS = 4;           // taps accumulated per CTA (one kernel row kh)
constexpr int WG_THREADS = 192;
constexpr int WG_BOX_BYTES = WG_P * 128;   // one [64 pixels x 64 channels] bf16 box

struct WgradParams {
  CUtensorMap tmS0, tmS1, tmG;
  int Wt, Ht, Bt, tiles_w, tiles_h, tiles_b;
  int M0, M1, N;
  int kblocks, kb_per_split;
  float* dw;
  // k3 = 1: 3x3 / stride 1 / pad 1 (binaural_attention_model.py DoubleConv): 3 rows of 3 taps, G at S's resolution.
  int k3;
  int ntaps;          // taps per CTA (= per kernel row): 4 or 3
  int taps_total;     // 16 or 9
  int ldn, n_off;     // dw row pitch per tap (total input channels) and column offset of this G tensor
  int g_pad;          // 4x4 mode: G carries an explicit one-pixel border (row 2i+kh of the padded tensor, never out of bounds)
  int skip_edge_rows; // 4x4 mode, Hs == 1 without a border: kernel rows 0 and 3 are all padding
};

template <int NT>
struct WgradSmem {
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;
  static constexpr int B_TAP_BYTES = (NT / 64) * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + WG_TAPS * B_TAP_BYTES;
  static constexpr int STAGES = NT == 128 ? 4 : 8;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = WG_TAPS * NT;   // 512 or 256
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int NT, bool K3>
__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using S = WgradSmem<NT>;
  constexpr int NTAPS = K3 ? 3 : 4;               // taps per CTA = one kernel row (compile time: the tap loops unroll)
  constexpr int STAGES = S::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * WG_M;
  const int n0 = blockIdx.y * NT;
  const int kh = blockIdx.z % NTAPS;             // tap group = kernel row
  // 4x4 stride-2 window over a two-row G (the 1 x 1 bottleneck): kernel rows 0 and 3 only meet the zero padding, their
  // gradient is exactly zero and dw is zeroed by the caller
  if (!K3 && p.skip_edge_rows && (kh == 0 || kh == 3)) return;
  const int split = blockIdx.z / NTAPS;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.kblocks);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmS0);
    if (p.M1 > 0) prefetch_tmap(&p.tmS1);
    prefetch_tmap(&p.tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const bool second = m0 >= p.M0;
      const CUtensorMap* tmS = second ? &p.tmS1 : &p.tmS0;
      const int mc = second ? m0 - p.M0 : m0;
      const int di = p.g_pad ? (kh >> 1) : (kh + 1) / 2 - 1, ra = p.g_pad ? (kh & 1) : (kh + 1) & 1;
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < nkb; ++it) {
        const int kb = kb_begin + it;
        const int tw_i = kb % p.tiles_w, th_i = (kb / p.tiles_w) % p.tiles_h, tb_i = kb / (p.tiles_w * p.tiles_h);
        const int x0 = tw_i * p.Wt, y0 = th_i * p.Ht, b0 = tb_i * p.Bt;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* a_dst = smem + s * S::STAGE_BYTES;
        unsigned char* b_dst = a_dst + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::A_BYTES + NTAPS * S::B_TAP_BYTES);
        tma_load_4d(a_dst, tmS, &full_bar[s], mc, x0, y0, b0);
        tma_load_4d(a_dst + WG_BOX_BYTES, tmS, &full_bar[s], mc + 64, x0, y0, b0);   // (rows past M: zero-filled)
#pragma unroll
        for (int kw = 0; kw < NTAPS; ++kw) {
          const int dj = p.g_pad ? (kw >> 1) : (kw + 1) / 2 - 1, rb = p.g_pad ? (kw & 1) : (kw + 1) & 1;
#pragma unroll
          for (int h = 0; h < NT / 64; ++h) {
            unsigned char* dst = b_dst + kw * S::B_TAP_BYTES + h * WG_BOX_BYTES;
            if (K3) tma_load_4d(dst, &p.tmG, &full_bar[s], n0 + h * 64, x0 + kw - 1, y0 + kh - 1, b0);
            else tma_load_5d(dst, &p.tmG, &full_bar[s], rb * p.N + n0 + h * 64, x0 + dj, ra, y0 + di, b0);
          }
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(WG_M, NT, 1, 1);
      // descriptors differ between stages / taps / K steps only in the start-address field: build once, then add
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t a_desc0 = umma_smem_desc(smem_base, WG_BOX_BYTES, 1024);
      const uint64_t b_desc0 = umma_smem_desc(smem_base + S::A_BYTES, WG_BOX_BYTES, 1024);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < nkb; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t stage_off = (uint64_t)((uint32_t)(s * S::STAGE_BYTES) >> 4);
#pragma unroll
        for (int kw = 0; kw < NTAPS; ++kw) {
#pragma unroll
          for (int k = 0; k < WG_P / 16; ++k) {
            // 16 pixels = two 8-row swizzle atoms = 2048 bytes further down the tile
            const uint64_t ad = a_desc0 + stage_off + (uint64_t)((k * 2048) >> 4);
            const uint64_t bd = b_desc0 + stage_off + (uint64_t)((kw * S::B_TAP_BYTES + k * 2048) >> 4);
            umma_bf16(tmem_base + kw * NT, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(accum_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int kw = 0; kw < NTAPS; ++kw) {
      float* row = p.dw + ((size_t)m * (NTAPS * NTAPS) + kh * NTAPS + kw) * p.ldn + p.n_off + n0;
#pragma unroll 1
      for (int cc = 0; cc < NT; cc += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kw * NT + cc), v);
        if (nkb > 0 && m < p.M0 + p.M1) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) red_add_v4(row + cc + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ plain transposed-operand GEMM
// D[128][NT] += sum_rows A[row][0:128] * Bm[row][0:NT]   (both operands row-major = MN-major for the MMA).
// Used by the thin layers' weight gradients, whose operands are [pixels][channels] / [pixels][patch] matrices.
struct GemmTnParams {
  CUtensorMap tmA0, tmA1, tmB;
  int ca0, ca1;
  int kblocks, kb_per_split;
  float* D;
  int cb;            // first column of Bm used by this launch
  int ldd, m_valid;  // row pitch of D and number of valid rows (<= 128)
  int m_total;       // > 0: blockIdx.y / blockIdx.z select the [128 x NT] block of an [m_total x N] result
};

template <int NT>
struct GemmTnSmem {
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;
  static constexpr int B_BYTES = (NT / 64) * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 8;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1) tc_gemm_tn_kernel(const __grid_constant__ GemmTnParams p) {
  using S = GemmTnSmem<NT>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.x * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.kblocks);
  const int nkb = kb_end - kb_begin;
  const int mb = p.m_total > 0 ? (int)blockIdx.y : 0, nb = p.m_total > 0 ? (int)blockIdx.z : 0;
  const int ca0 = p.ca0 + mb * 128, ca1 = p.ca1 + mb * 128, cb = p.cb + nb * NT;
  const int m_valid = p.m_total > 0 ? min(128, p.m_total - mb * 128) : p.m_valid;
  float* const Dblk = p.D + (size_t)mb * 128 * p.ldd + (size_t)nb * NT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmA0);
    prefetch_tmap(&p.tmA1);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < nkb; ++it) {
        const int row0 = (kb_begin + it) * WG_P;
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* a_dst = smem + s * S::STAGE_BYTES;
        unsigned char* b_dst = a_dst + S::A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        tma_load_2d(a_dst, &p.tmA0, &full_bar[s], ca0, row0);
        tma_load_2d(a_dst + WG_BOX_BYTES, &p.tmA1, &full_bar[s], ca1, row0);
#pragma unroll
        for (int h = 0; h < NT / 64; ++h) tma_load_2d(b_dst + h * WG_BOX_BYTES, &p.tmB, &full_bar[s], cb + h * 64, row0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(WG_M, NT, 1, 1);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_P / 16; ++k) {
          const uint64_t ad = umma_smem_desc(a_addr + k * 2048, WG_BOX_BYTES, 1024);
          const uint64_t bd = umma_smem_desc(b_addr + k * 2048, WG_BOX_BYTES, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    float* row = Dblk + (size_t)m * p.ldd;
#pragma unroll 1
    for (int cc = 0; cc < NT; cc += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      if (nkb > 0 && m < m_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) red_add_v4(row + cc + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
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

template <int NT>
int launch_gemm_tn(const GemmTnParams& p, int splits, cudaStream_t s, int mblocks = 1, int nblocks = 1) {
  using S = GemmTnSmem<NT>;
  ADP_SMEM_ATTR(tc_gemm_tn_kernel<NT>, S::BYTES);
  tc_gemm_tn_kernel<NT><<<dim3(splits, mblocks, nblocks), WG_THREADS, S::BYTES, s>>>(p);
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

bool wg_geometry(int Hs, int Ws, int* Wt, int* Ht, int* Bt) {
  if (!pow2(Hs) || !pow2(Ws)) return false;
  *Wt = Ws < WG_P ? Ws : WG_P;
  int rest = WG_P / *Wt;
  *Ht = Hs < rest ? Hs : rest;
  *Bt = rest / *Ht;
  return true;
}

template <int NT, bool K3>
int launch_wgrad(const WgradParams& p, dim3 grid, cudaStream_t s) {
  using S = WgradSmem<NT>;
  ADP_SMEM_ATTR((tc_wgrad_kernel<NT, K3>), S::BYTES);
  tc_wgrad_kernel<NT, K3><<<grid, WG_THREADS, S::BYTES, s>>>(p);
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

}  // namespace

int tc_gemm_tn(const void* a0, int lda0, int ca0, const void* a1, int lda1, int ca1, const void* bm, int ldb, int NT,
               long long rows, float* D, cudaStream_t s) {
  ADP_CHECK_ARG(NT == 64 || NT == 128, "tc_gemm_tn: NT must be 64 or 128");
  ADP_CHECK_ARG(lda0 % 8 == 0 && lda1 % 8 == 0 && ldb % 8 == 0 && ldb >= NT && rows > 0 && rows < (1LL << 31),
                "tc_gemm_tn: bad leading dimensions / rows");
  GemmTnParams p;
  memset(&p, 0, sizeof(p));
  const void* bases[3] = {a0, a1, bm};
  const int lds[3] = {lda0, lda1, ldb};
  CUtensorMap* maps[3] = {&p.tmA0, &p.tmA1, &p.tmB};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[2] = {(uint64_t)lds[i], (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)lds[i] * 2};
    uint32_t box[2] = {64, (uint32_t)WG_P};
    ADP_TRY(tc::make_tmap_bf16(maps[i], bases[i], 2, dims, str, box));
  }
  p.ca0 = ca0; p.ca1 = ca1; p.D = D;
  p.cb = 0; p.ldd = NT; p.m_valid = 128;
  p.kblocks = (int)((rows + WG_P - 1) / WG_P);
  int splits = 2 * sm_count();
  if (splits > p.kblocks) splits = p.kblocks;
  p.kb_per_split = adp_cdiv(p.kblocks, splits);
  splits = adp_cdiv(p.kblocks, p.kb_per_split);
  ADP_CUDA(cudaMemsetAsync(D, 0, sizeof(float) * 128 * NT, s));
  if (NT == 128) return launch_gemm_tn<128>(p, splits, s);
  return launch_gemm_tn<64>(p, splits, s);
}

// dw[m][n] += sum_rows A[row][m] * Bm[row][n]  for an [M x N] fp32 matrix (row pitch ldd), M % 64 == 0, N % 64 == 0:
// the weight gradient of a 1x1 convolution (A = dL/dy [pixels][Cout], Bm = layer input [pixels][Cin]) and the
// transposed products of the attention backward (dV = P^T dO, dK = dS^T Q).  One launch, one CTA per ([128 x NT] block,
// row split); the caller zeroes dw.
int tc_gemm_tn_full(const void* a, int M, const void* bm, int N, float* dw, int ldd, long long rows, cudaStream_t s) {
  ADP_CHECK_ARG(M > 0 && M % 64 == 0 && N > 0 && N % 64 == 0 && rows > 0 && rows < (1LL << 31) && ldd >= N,
                "tc_gemm_tn_full: unsupported shape M=%d N=%d rows=%lld", M, N, rows);
  GemmTnParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {(uint64_t)M, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)M * 2};
    uint32_t box[2] = {64, (uint32_t)WG_P};
    ADP_TRY(tc::make_tmap_bf16(&p.tmA0, a, 2, dims, str, box));
    p.tmA1 = p.tmA0;
  }
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, (uint32_t)WG_P};
    ADP_TRY(tc::make_tmap_bf16(&p.tmB, bm, 2, dims, str, box));
  }
  const int NT = N % 128 == 0 ? 128 : 64;
  p.kblocks = (int)((rows + WG_P - 1) / WG_P);
  const int mblocks = adp_cdiv(M, 128), nblocks = N / NT;
  ADP_CHECK_ARG(mblocks <= 65535 && nblocks <= 65535, "tc_gemm_tn_full: result too large");
  int splits = adp_cdiv(2 * sm_count(), mblocks * nblocks);
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.kb_per_split = adp_cdiv(p.kblocks, splits);
  splits = adp_cdiv(p.kblocks, p.kb_per_split);
  p.ldd = ldd; p.D = dw; p.m_total = M;
  p.ca0 = 0; p.ca1 = 64; p.cb = 0;                 // (a 64-column half past M is outside the tensor: zero-filled, not stored)
  if (NT == 128) return launch_gemm_tn<128>(p, splits, s, mblocks, nblocks);
  return launch_gemm_tn<64>(p, splits, s, mblocks, nblocks);
}

bool tc_supported_wgrad(int B, int Hs, int Ws, int M0, int M1, int N) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !tc::encode_tiled_fn()) return false;
  if (M0 <= 0 || M0 % WG_M || M1 % WG_M || N % 64 || B < 1) return false;
  return wg_geometry(Hs, Ws, &Wt, &Ht, &Bt);
}

int tc_wgrad(const void* s0, int M0, const void* s1, int M1, const void* g, int N, float* dw, int B, int Hs, int Ws,
             cudaStream_t s, int g_pad) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(wg_geometry(Hs, Ws, &p.Wt, &p.Ht, &p.Bt), "tc_wgrad: unsupported spatial size %dx%d", Hs, Ws);
  ADP_CHECK_ARG(M0 % WG_M == 0 && M1 % WG_M == 0 && N % 64 == 0, "tc_wgrad: unsupported channels");
  const int NT = N % 128 == 0 ? 128 : 64;
  p.tiles_w = Ws / p.Wt; p.tiles_h = Hs / p.Ht; p.tiles_b = adp_cdiv(B, p.Bt);
  p.M0 = M0; p.M1 = M1; p.N = N; p.dw = dw;
  p.k3 = 0; p.ntaps = 4; p.taps_total = 16; p.ldn = N; p.n_off = 0;
  p.kblocks = p.tiles_w * p.tiles_h * p.tiles_b;
  for (int h = 0; h < 2; ++h) {
    const int C = h == 0 ? M0 : M1;
    if (C == 0) continue;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Ws * C * 2, (uint64_t)Hs * Ws * C * 2};
    uint32_t box[4] = {64, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmS0 : &p.tmS1, h == 0 ? s0 : s1, 4, dims, str, box));
  }
  p.g_pad = g_pad ? 1 : 0;
  p.skip_edge_rows = (Hs == 1 && !p.g_pad) ? 1 : 0;
  {  // G [B, 2Hs, 2Ws, N] viewed as (2N | Ws | 2 | Hs | B); with a border: [B, 2Hs+2, 2Ws+2, N] as (2N | Ws+1 | 2 | Hs+1 | B)
    const int Hg = 2 * Hs + 2 * p.g_pad, Wg = 2 * Ws + 2 * p.g_pad;
    uint64_t dims[5] = {(uint64_t)2 * N, (uint64_t)Wg / 2, 2, (uint64_t)Hg / 2, (uint64_t)B};
    uint64_t str[4] = {(uint64_t)2 * N * 2, (uint64_t)Wg * N * 2, (uint64_t)2 * Wg * N * 2, (uint64_t)Hg * Wg * N * 2};
    uint32_t box[5] = {64, (uint32_t)p.Wt, 1, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(&p.tmG, g, 5, dims, str, box));
  }
  const int m_tiles = (M0 + M1) / WG_M, n_tiles = N / NT;
  const long long ctas = (long long)m_tiles * n_tiles * 4;
  // one resident CTA per SM (160 KB of smem, all 512 TMEM columns): aim at ADP_WG_WAVES (default 1) full waves
  static int waves = getenv("ADP_WG_WAVES") ? atoi(getenv("ADP_WG_WAVES")) : 1;
  const int eff_waves = NT == 64 ? 2 * waves : waves;      // the narrow-N tiles are short: two waves balance better
  // (rounded down: 152 or 160 CTAs on 148 SMs cost a second, nearly empty wave -- E3 / E4 at B = 64)
  int splits = (int)(((long long)eff_waves * sm_count()) / ctas);
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.kb_per_split = adp_cdiv(p.kblocks, splits);
  splits = adp_cdiv(p.kblocks, p.kb_per_split);
  dim3 grid(m_tiles, n_tiles, 4 * splits);
  if (NT == 128) return launch_wgrad<128, false>(p, grid, s);
  return launch_wgrad<64, false>(p, grid, s);
}

// Weight gradient of the 3x3 / stride 1 / pad 1 convolution:
//   dw[m][kh,kw][n_off + n] += sum_{b,i,j} S[b,i,j,m] * G[b,i+kh-1,j+kw-1,n]     (S = dL/dy [M], G = layer input [N])
// dw is fp32 [M][9][ldn]; a concatenated input is handled by one call per half (n_off = 0 / C0).  M % 64 == 0
// (a 64-row tail is zero-filled by TMA and not stored), N % 64 == 0.
bool tc_supported_wgrad3x3(int B, int H, int W, int M, int N) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !tc::encode_tiled_fn()) return false;
  if (M <= 0 || M % 64 || N % 64 || N <= 0 || B < 1) return false;
  return wg_geometry(H, W, &Wt, &Ht, &Bt);
}

int tc_wgrad3x3(const void* sgrad, int M, const void* g, int N, int ldn, int n_off, float* dw, int B, int H, int W,
                cudaStream_t s) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(wg_geometry(H, W, &p.Wt, &p.Ht, &p.Bt), "tc_wgrad3x3: unsupported spatial size %dx%d", H, W);
  ADP_CHECK_ARG(M % 64 == 0 && N % 64 == 0 && ldn >= n_off + N && ldn % 4 == 0 && n_off % 4 == 0, "tc_wgrad3x3: unsupported channels");
  const int NT = N % 128 == 0 ? 128 : 64;
  p.tiles_w = W / p.Wt; p.tiles_h = H / p.Ht; p.tiles_b = adp_cdiv(B, p.Bt);
  p.M0 = M; p.M1 = 0; p.N = N; p.dw = dw;
  p.k3 = 1; p.ntaps = 3; p.taps_total = 9; p.ldn = ldn; p.n_off = n_off;
  p.kblocks = p.tiles_w * p.tiles_h * p.tiles_b;
  {
    uint64_t dims[4] = {(uint64_t)M, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)M * 2, (uint64_t)W * M * 2, (uint64_t)H * W * M * 2};
    uint32_t box[4] = {64, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(&p.tmS0, sgrad, 4, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)N * 2, (uint64_t)W * N * 2, (uint64_t)H * W * N * 2};
    uint32_t box[4] = {64, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(&p.tmG, g, 4, dims, str, box));
  }
  const int m_tiles = adp_cdiv(M, WG_M), n_tiles = N / NT;
  const long long ctas = (long long)m_tiles * n_tiles * 3;
  int splits = (int)(((long long)(NT == 64 ? 2 : 1) * sm_count()) / ctas);
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.kb_per_split = adp_cdiv(p.kblocks, splits);
  splits = adp_cdiv(p.kblocks, p.kb_per_split);
  dim3 grid(m_tiles, n_tiles, 3 * splits);
  if (NT == 128) return launch_wgrad<128, true>(p, grid, s);
  return launch_wgrad<64, true>(p, grid, s);
}

}  // namespace adp
