// tcgen05 / TMEM / TMA implicit-GEMM convolutions for the U-Net hidden layers (bf16 operands, fp32
// accumulation in tensor memory).  Replaces cuDNN's kernels behind nn.Conv2d(k4,s2,p1) and
// nn.ConvTranspose2d(k4,s2,p1) (models/unetbaseline_model.py:187,:196,:209,:218).
//
// Kernel 1 (tc_igemm_persist_kernel) serves the "gather" (F1) and "parity" (F2) families of
// adp_conv_simt.cu as   D[128 pixels x BLOCK_N channels] = sum_kblocks A[128 x 64] * W[BLOCK_N x 64]^T :
//   * A tiles are fetched by TMA straight from the NHWC activation tensor.  No im2col, no padded or
//     space-to-depth copy: the stride-2 4x4 window is expressed through a 5-D view of the tensor
//     (2C | W/2 | row parity | H/2 | B) whose boxes, shifted by one per tap, cover exactly the
//     128 output pixels of the tile; out-of-bounds box elements (the conv padding, and the batch
//     tail) are zero-filled by the TMA unit.  The transposed conv's parity classes use a 4-D view.
//     The decoder's skip concat is two tensor maps: the K loop switches map at C0.
//   * W tiles come from the bf16 weight operand [N][16][C] (K-major), 2-D map.
//   * one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) on SWIZZLE_128B smem
//     descriptors; completion is signalled to the TMA producer / epilogue through tcgen05.commit
//     on mbarriers; a STAGES-deep smem ring overlaps TMA with MMA.
//   * the epilogue (4 warps = 128 TMEM lanes) reads the accumulator with tcgen05.ld and writes bf16
//     NHWC rows (strided by parity for F2, split into two tensors for a concat gradient), or
//     accumulates fp32 partial sums when the K range is split across CTAs (deep, small-M layers); those are
//     finished (rounded to bf16, cleared) by finish_partial_kernel or, inside the U-Net engine, by the
//     BatchNorm / activation kernel that consumes the layer's output (ConvExtras::deferred).
#include <stdlib.h>
#include "adp_tc.cuh"

namespace adp {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    adp_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ADP_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    adp_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u]", (int)r, rank,
                  (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0), (unsigned long long)(rank > 2 ? gd[2] : 0),
                  (unsigned long long)(rank > 3 ? gd[3] : 0), (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0],
                  rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return ADP_ERR_CUDA;
  }
  return ADP_OK;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int TILE_M = 128;
constexpr int TILE_K = 64;            // bf16 elements = one 128-byte swizzled row
constexpr int A_STAGE_BYTES = TILE_M * TILE_K * 2;
// threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocation, then 4 (convolutions) or 8 (STFT) epilogue warps

struct IgemmParams {
  CUtensorMap tmA0, tmA1, tmW;
  CUtensorMap tmA2;                   // mode 3: third bf16 slice of the frames
  // tile geometry over the "small" pixel grid (conv outputs for F1, convT inputs for F2)
  int Wt, Ht, Bt, tiles_w, tiles_h;
  int B, Hs, Ws;                      // small grid extent
  int mode;                           // 0 = gather (F1), 1 = parity (F2), 2 = pointwise (1x1), 3 = STFT, 4 = 3x3 stride 1 pad 1
  int C0, C1, Ct;                     // input channels (concat halves)
  int N, N0, N1;                      // output channels and split
  int kblocks, splits, kb_per_split;
  bf16* y0; bf16* y1;
  float* partial;                     // fp32 [out pixels][N] when splits > 1
  float* out_f32;                     // mode 2: fp32 [pixels][BLOCK_N] result
  int gx, gy, gz, total_tiles;        // logical grid (x fastest) walked by the persistent kernel
  // Attention epilogues of the row GEMM (mode 2): D = A * B^T is a tile of the score matrix (row = opix, column = n)
  //   epi 1: stat_m[row] = max(stat_m[row], max_n scale*D)                     (ordered-int atomicMax)
  //   epi 2: stat_l[row] += sum_n exp(scale*D - m[row])
  //   epi 3: out bf16 = exp(scale*D - m[i]) / l[i]            i = row, or i = column when stat_by_col (transposed scores)
  //   epi 4: out bf16 = scale * pmat[row][n] * (D - delta[i])                  (softmax backward, D = dP)
  int epi, stat_by_col;
  float escale;
  int* stat_m; float* stat_l; const float* delta; const bf16* pmat;
  // 1 x 1 bottleneck (E8 / D8 at the innermost level): a 4x4 stride-2 window over a 2 x 2 input only ever sees taps
  // (1..2, 1..2), a transposed conv from one input pixel uses one tap per output parity -- the other 12 taps multiply the
  // zero padding, so their k-blocks (75 % of the weight bytes) are skipped
  unsigned tap_lut;                   // mode 0: nibble i = i-th tap to visit (0 = all 16 in order)
  int one_tap;                        // mode 1: the single valid tap (th, tw) = (1 - pa, 1 - pb) per parity class
  int halo;                           // 0, or the HALO instantiation: 4 = parity mode, 3 = 3x3 rows
  int wmode;                          // modes 2/4: 1 = weights through the MN-major 3-D map (N | taps | Ct), taps reversed
  int f32_rows;                       // modes 2/4: write fp32 [pixels][N] to out_f32 instead of bf16 (any BLOCK_N)
  long long rows_guard;               // > 0: output rows (opix) >= rows_guard are not stored (ragged GEMM M)
  int act_dual;                       // y0 = lrelu(z, slope0) [- center], y1 (optional) = lrelu(z, slope1), both [pixels][N];
  float slope0, slope1;               //   z = D, or D * aff_scale[n] + aff_shift[n] (eval-mode BatchNorm folded in)
  const float* aff_scale; const float* aff_shift;
  // mode 3 (STFT as a split-bf16 DFT GEMM): rows = frames, columns = (re, im) pairs of the bins
  float* spec; int F, T, log_mode; int* minmax;
  // first-level centring (adp_unet.cu): the activation with the large per-channel DC component is stored as a - center[n]
  // inside a tensor with an explicit one-pixel border holding -center[n], so that the next convolution sees exact zeros
  // where the reference pads (bf16 keeps its 8 mantissa bits for the signal instead of the DC)
  int pad_in;                         // mode 0: the input is [B, Hi+2, Wi+2, C] with that border; no tap is out of bounds
  int pad_out;                        // act_dual: y0 is the interior of a [B, Hs+2, Ws+2, N] tensor
  const float* center;                // act_dual: subtracted from y0 after the activation (NULL: nothing)
  // BatchNorm statistics of the stored (bf16-rounded) output: stats[n] += sum, stats[N + n] += sum of squares (staged
  // bf16 epilogue only; NULL: not wanted)
  double* stats;
};

// ------------------------------------------------------------------ persistent implicit-GEMM kernel
// One CTA per SM walks the tile list (tile = blockIdx.x + i * gridDim.x).  The smem ring and its phases run
// continuously across tiles, the accumulator is double-buffered in tensor memory (2 x BLOCK_N columns), so the
// epilogue of tile i (TMEM -> registers -> global) overlaps the TMA/MMA main loop of tile i+1, and TMEM allocation,
// barrier initialisation and tensor-map prefetch are paid once per SM instead of once per tile.
struct TileCoord { int x0, y0c, b0, n0, pa, pb, kb_begin, nkb; };

template <int BLOCK_N>
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int t) {
  const int bx = t % p.gx, r = t / p.gx, by = r % p.gy, bz = r / p.gy;
  TileCoord c;
  // parity mode: the four output-parity classes of one pixel tile read the same input pixels (different taps);
  // they are adjacent in the walk so that the tile is fetched from HBM once and hit in L2 three times
  const int tm = p.mode == 1 ? (bx >> 2) : bx;
  const int tw_i = tm % p.tiles_w, th_i = (tm / p.tiles_w) % p.tiles_h, tb_i = tm / (p.tiles_w * p.tiles_h);
  c.x0 = tw_i * p.Wt; c.y0c = th_i * p.Ht; c.b0 = tb_i * p.Bt;
  c.n0 = by * BLOCK_N;
  const int zpar = p.mode == 1 ? (bx & 3) : 0;
  c.pa = zpar >> 1; c.pb = zpar & 1;
  c.kb_begin = bz * p.kb_per_split;
  c.nkb = min(c.kb_begin + p.kb_per_split, p.kblocks) - c.kb_begin;
  return c;
}

// HALO (BLOCK_N <= 128): the input window of a 16 x 8 pixel tile is loaded ONCE per k-block and several filter taps read it
// through shifted descriptors (the SWIZZLE_128B pattern is a function of the shared-memory address, so an operand may
// start at any 128-byte row and use any group pitch: tools/probe_umma_offset.py), next to those taps' weight tiles.
//   HALO = 4: parity (transposed-conv) mode, k-block = one 64-channel chunk, 2 x 2 taps, window 17 x 9 pixels
//   HALO = 3: 3x3 / stride 1 mode, k-block = (kernel row, chunk), 3 taps of that row, window 16 x 10 pixels
template <int HALO> struct HaloGeom {
  static constexpr int W = HALO == 3 ? 10 : 9, H = HALO == 3 ? 16 : 17;
  static constexpr int BOX_BYTES = W * H * TILE_K * 2;            // 20480 / 19584
  static constexpr int TAPS = HALO == 3 ? 3 : 4;
};
// ALT: two epilogue warp groups that take alternate TILES (one TMEM accumulator buffer each) -- the thin pointwise layers
// have a 4-MMA main loop and are bound by the latency of one epilogue group; no fused statistics in that variant
template <int BLOCK_N, int HALO = 0, bool ALT = false>
struct PersistSmem {
  static constexpr int A_BYTES = HALO ? 20480 : A_STAGE_BYTES;
  static constexpr int B_TILE = BLOCK_N * TILE_K * 2;
  static constexpr int B_BYTES = HALO ? HaloGeom<HALO>::TAPS * B_TILE : B_TILE;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING = (ALT ? 8 : 4) * 4096;   // epilogue transpose buffers: one (32 rows x 128 B) per epilogue warp
  static constexpr int BAR_BYTES = 512;           // pipeline barriers, TMEM slot
  // BatchNorm partial sums, private to each of the 4 epilogue warps: [4][2][STATS_N] floats.  The 64-wide parity halo
  // kernel (N = 64 only) keeps them in the 896 unused bytes behind each stage's 17 x 9 window instead, so that its
  // 4th pipeline stage survives; the 3x3 halo kernels (config 4) never fuse statistics.
  static constexpr bool STATS_IN_SLACK = HALO == 4 && BLOCK_N == 64;
  static constexpr int STATS_N = (HALO == 3 || ALT) ? 0 : (STATS_IN_SLACK ? 64 : 512);
  static constexpr int STATS_BYTES = STATS_IN_SLACK ? 0 : 4 * 2 * STATS_N * 4;
  static constexpr int RING_BUDGET = HALO ? (227 * 1024 - 1024 - BAR_BYTES - STAGING - STATS_BYTES - 256) : 192 * 1024;
  static constexpr int STAGES = RING_BUDGET / STAGE_BYTES > 10 ? 10 : RING_BUDGET / STAGE_BYTES;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + STAGING + STATS_BYTES;
  static constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  static_assert(BYTES <= 227 * 1024, "shared-memory budget");
  static_assert(!STATS_IN_SLACK || (STAGES >= 4 && A_BYTES - HaloGeom<HALO ? HALO : 4>::BOX_BYTES >= 2 * STATS_N * 4), "statistics slack");
};

// EG = number of epilogue warp groups (4 warps each): 1 for the convolutions, 2 for the math-heavy STFT epilogue
// ATT = attention (softmax) epilogues of the row GEMM compiled in (kept out of the convolution instantiations)
template <int BLOCK_N, int EG, bool ATT = false, int HALO = 0, bool ALT = false>
__global__ void __launch_bounds__(64 + 128 * EG, 1) tc_igemm_persist_kernel(const __grid_constant__ IgemmParams p) {
  static_assert(!ALT || (EG == 2 && HALO == 0 && !ATT), "alternating epilogue groups: two groups, plain staged epilogue");
  using PS = PersistSmem<BLOCK_N, HALO, ALT>;
  const int worker = (int)blockIdx.x;                     // index of this CTA in the tile walk
  const int nworkers = (int)gridDim.x;
  const int ntiles = p.total_tiles;
  constexpr int STAGES = PS::STAGES;
  constexpr int ACC = PS::ACC_COLS;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * PS::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained by the 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool want_stats = EG == 1 && !ALT && BLOCK_N >= 64 && PS::STATS_N > 0 && p.stats != nullptr;
  // BN partial sums of epilogue warp e: [2][N] floats
  auto stats_of = [&](int e) -> float* {
    if (PS::STATS_IN_SLACK) return reinterpret_cast<float*>(smem + e * PS::STAGE_BYTES + HaloGeom<HALO ? HALO : 4>::BOX_BYTES);
    return reinterpret_cast<float*>(smem + STAGES * PS::STAGE_BYTES + PS::BAR_BYTES + PS::STAGING) + e * 2 * p.N;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmA0);
    if (p.C1 > 0) prefetch_tmap(&p.tmA1);
    prefetch_tmap(&p.tmW);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], ALT ? 4 : 4 * EG); }
    fence_barrier_init();
  }
  if (want_stats) {
    for (int e = 0; e < 4; ++e)
      for (int i = threadIdx.x; i < 2 * p.N; i += blockDim.x) stats_of(e)[i] = 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * ACC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int nchunk = p.Ct / TILE_K;
      int s = 0;                                          // ring position and its phase bit
      uint32_t ph = 0;
      for (int t = worker; t < ntiles; t += nworkers) {
        const TileCoord c = decode_tile<BLOCK_N>(p, t);
        int tap = c.kb_begin / nchunk, ch = (c.kb_begin - tap * nchunk) * TILE_K;
        for (int it = 0; it < c.nkb; ++it) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* a_dst = smem + s * PS::STAGE_BYTES;
          unsigned char* b_dst = a_dst + PS::A_BYTES;
          mbar_expect_tx(&full_bar[s], HALO ? HaloGeom<HALO>::BOX_BYTES + HaloGeom<HALO>::TAPS * PS::B_TILE : PS::STAGE_BYTES);
          if (HALO == 4) {
            // one 64-channel chunk: the tile's halo window once + the weight tiles of the four taps of this parity
            const int cx = c.x0 + c.pb - 1, cy = c.y0c + c.pa - 1;
            if (ch < p.C0) tma_load_4d(a_dst, &p.tmA0, &full_bar[s], ch, cx, cy, c.b0);
            else tma_load_4d(a_dst, &p.tmA1, &full_bar[s], ch - p.C0, cx, cy, c.b0);
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              const int wtap = (3 - c.pa - 2 * (t4 >> 1)) * 4 + (3 - c.pb - 2 * (t4 & 1));
#pragma unroll
              for (int h = 0; h < (BLOCK_N >= 64 ? BLOCK_N / 64 : 1); ++h)
                tma_load_3d(b_dst + t4 * PS::B_TILE + h * (TILE_K * 128), &p.tmW, &full_bar[s], c.n0 + h * 64, wtap, ch);
            }
          } else if (HALO == 3) {
            // `tap` counts kernel rows here: rows y0+kh-1 .. +15, columns x0-1 .. x0+8, and the three taps of that row
            const int kh = tap;
            if (ch < p.C0) tma_load_4d(a_dst, &p.tmA0, &full_bar[s], ch, c.x0 - 1, c.y0c + kh - 1, c.b0);
            else tma_load_4d(a_dst, &p.tmA1, &full_bar[s], ch - p.C0, c.x0 - 1, c.y0c + kh - 1, c.b0);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int t9 = kh * 3 + kw;
              if (p.wmode) {
#pragma unroll
                for (int h = 0; h < (BLOCK_N >= 64 ? BLOCK_N / 64 : 1); ++h)
                  tma_load_3d(b_dst + kw * PS::B_TILE + h * (TILE_K * 128), &p.tmW, &full_bar[s], c.n0 + h * 64, 8 - t9, ch);
              } else {
                tma_load_2d(b_dst + kw * PS::B_TILE, &p.tmW, &full_bar[s], t9 * p.Ct + ch, c.n0);
              }
            }
          } else if (p.mode == 0) {
            // 4x4 stride-2 window over the (2C | W/2 | row parity | H/2 | B) view: tap row kh -> (row offset, row parity)
            // = kh:0 (-1,1) 1 (0,0) 2 (0,1) 3 (+1,0); with an explicit border (pad_in) row 2oy+kh -> (kh >> 1, kh & 1)
            const int atap = p.tap_lut ? (int)((p.tap_lut >> (4 * tap)) & 15u) : tap;
            const int kh = atap >> 2, kw = atap & 3;
            const int di = p.pad_in ? (kh >> 1) : (kh + 1) / 2 - 1, ra = p.pad_in ? (kh & 1) : (kh + 1) & 1;
            const int dj = p.pad_in ? (kw >> 1) : (kw + 1) / 2 - 1, rb = p.pad_in ? (kw & 1) : (kw + 1) & 1;
            tma_load_5d(a_dst, &p.tmA0, &full_bar[s], rb * p.Ct + ch, c.x0 + dj, ra, c.y0c + di, c.b0);
            tma_load_2d(b_dst, &p.tmW, &full_bar[s], atap * p.Ct + ch, c.n0);
          } else if (p.mode == 1) {
            const int th = p.one_tap ? 1 - c.pa : tap >> 1, tw = p.one_tap ? 1 - c.pb : tap & 1;
            const int cx = c.x0 + c.pb - 1 + tw, cy = c.y0c + c.pa - 1 + th;
            if (ch < p.C0) tma_load_4d(a_dst, &p.tmA0, &full_bar[s], ch, cx, cy, c.b0);
            else tma_load_4d(a_dst, &p.tmA1, &full_bar[s], ch - p.C0, cx, cy, c.b0);
            // weights stay in their master layout [c][tap][n]: the B tile is fetched N-major (64-row boxes of
            // 64 output channels) and handed to the MMA through an MN-major descriptor -- no transposed copy
            const int wtap = (3 - c.pa - 2 * th) * 4 + (3 - c.pb - 2 * tw);
#pragma unroll
            for (int h = 0; h < (BLOCK_N >= 64 ? BLOCK_N / 64 : 1); ++h)
              tma_load_3d(b_dst + h * (TILE_K * 128), &p.tmW, &full_bar[s], c.n0 + h * 64, wtap, ch);
          } else if (p.mode == 3) {
            // six K blocks x0*W0, x0*W1, x1*W0, x0*W2, x1*W1, x2*W0 of the three-way bf16 split (A slice 0,0,1,0,1,2)
            const int kbi = ch / TILE_K;
            const CUtensorMap* am = (kbi == 2 || kbi == 4) ? &p.tmA1 : (kbi == 5 ? &p.tmA2 : &p.tmA0);
            tma_load_4d(a_dst, am, &full_bar[s], 0, c.x0, c.y0c, c.b0);
            tma_load_2d(b_dst, &p.tmW, &full_bar[s], ch, c.n0);
          } else if (p.mode == 4 || (p.mode == 2 && p.wmode)) {
            // 3x3 / stride 1 / pad 1 (or a 1x1 with MN-major weights): tap (th, tw) reads the input shifted by (th-1, tw-1)
            const int th = p.mode == 4 ? tap / 3 : 1, tw = p.mode == 4 ? tap - 3 * th : 1;
            const int cx = c.x0 + tw - 1, cy = c.y0c + th - 1;
            if (ch < p.C0) tma_load_4d(a_dst, &p.tmA0, &full_bar[s], ch, cx, cy, c.b0);
            else tma_load_4d(a_dst, &p.tmA1, &full_bar[s], ch - p.C0, cx, cy, c.b0);
            if (p.wmode) {   // data gradient: w[k = cout][8 - tap][n = cin], n contiguous
              const int wtap = p.mode == 4 ? 8 - tap : 0;
#pragma unroll
              for (int h = 0; h < (BLOCK_N >= 64 ? BLOCK_N / 64 : 1); ++h)
                tma_load_3d(b_dst + h * (TILE_K * 128), &p.tmW, &full_bar[s], c.n0 + h * 64, wtap, ch);
            } else {
              tma_load_2d(b_dst, &p.tmW, &full_bar[s], tap * p.Ct + ch, c.n0);
            }
          } else if (p.mode == 2) {
            if (ch < p.C0) tma_load_4d(a_dst, &p.tmA0, &full_bar[s], ch, c.x0, c.y0c, c.b0);
            else tma_load_4d(a_dst, &p.tmA1, &full_bar[s], ch - p.C0, c.x0, c.y0c, c.b0);
            tma_load_2d(b_dst, &p.tmW, &full_bar[s], ch, c.n0);
          }
          ch += TILE_K;
          if (ch == p.Ct) { ch = 0; ++tap; }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const bool b_mn = p.mode == 1 || p.wmode != 0;
      const uint32_t idesc = umma_idesc_bf16(TILE_M, BLOCK_N, 0, b_mn ? 1 : 0);
      // descriptors differ from stage to stage only in the 14-bit start-address field: build them once
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t a_desc0 = umma_smem_desc(smem_base, 16, HALO ? HaloGeom<HALO>::W * 128 : 1024);   // HALO: 8-pixel groups one window row apart
      const uint64_t b_desc0 = b_mn ? umma_smem_desc(smem_base + PS::A_BYTES, TILE_K * 128, 1024)
                                    : umma_smem_desc(smem_base + PS::A_BYTES, 16, 1024);
      const uint32_t b_kstep = b_mn ? (2048u >> 4) : (32u >> 4);
      int s = 0;
      uint32_t ph = 0, local = 0;
      for (int t = worker; t < ntiles; t += nworkers, ++local) {
        const TileCoord c = decode_tile<BLOCK_N>(p, t);
        const uint32_t buf = local & 1u, use = local >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1u) ^ 1u);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * ACC;
        for (int it = 0; it < c.nkb; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t stage_off = (uint64_t)((uint32_t)(s * PS::STAGE_BYTES) >> 4);
          const uint64_t ad0 = a_desc0 + stage_off, bd0 = b_desc0 + stage_off;
          if (HALO) {
            // (not unrolled over the taps: the fully unrolled form keeps 32 descriptors = 64+ uniform registers live and
            // was observed to raise "illegal instruction" on the UTCHMMA sequence depending on the surrounding code)
#pragma unroll 1
            for (int t4 = 0; t4 < HaloGeom<HALO>::TAPS; ++t4) {
              // window shifted by (th, tw) [parity] or by tw [3x3 row] pixels
              const int shift_rows = HALO != 3 ? (t4 >> 1) * HaloGeom<HALO>::W + (t4 & 1) : t4;
              const uint64_t bt4 = bd0 + (uint64_t)((t4 * PS::B_TILE) >> 4);
              const uint64_t at = ad0 + (uint64_t)((shift_rows * 128) >> 4);
#pragma unroll
              for (int k = 0; k < TILE_K / 16; ++k)
                umma_bf16(tacc, at + (uint64_t)(k * 2), bt4 + (uint64_t)(k * b_kstep), idesc, (it | t4 | k) != 0 ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < TILE_K / 16; ++k)
              umma_bf16(tacc, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * b_kstep), idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ===================== epilogue =====================
    // 8 warps: warp & 3 selects the TMEM lane quarter (hardware rule), (warp - 2) / 4 the group; the two groups
    // take alternate 32-column chunks so that twice as many warps hide the tcgen05.ld / math / store latency
    const int q = warp & 3;
    const int egroup = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int wt = r % p.Wt, ht = (r / p.Wt) % p.Ht, bt = r / (p.Wt * p.Ht);
    constexpr bool STAGED_OK = BLOCK_N >= 64 && (EG == 1 || ALT);
    unsigned char* stg = smem + STAGES * PS::STAGE_BYTES + PS::BAR_BYTES + (STAGED_OK ? (warp - 2) * 4096 : 0);
    const bool staged = p.splits <= 1 && p.mode != 3 && !p.f32_rows && !(ATT && (p.epi == 1 || p.epi == 2)) &&
                        (p.act_dual || p.N1 == 0 || p.N0 % 64 == 0);
    uint32_t local = 0;
    for (int t = worker; t < ntiles; t += nworkers, ++local) {
      if (ALT && (int)(local & 1u) != egroup) continue;     // the other group's tile (and accumulator buffer)
      const TileCoord c = decode_tile<BLOCK_N>(p, t);
      const uint32_t buf = local & 1u, use = local >> 1;
      mbar_wait(&tfull_bar[buf], use & 1u);
      tc_fence_after();
      const int b = c.b0 + bt, py = c.y0c + ht, px = c.x0 + wt;
      size_t opix;
      if (p.mode != 1) opix = ((size_t)b * p.Hs + py) * p.Ws + px;
      else opix = ((size_t)b * 2 * p.Hs + 2 * py + c.pa) * (2 * p.Ws) + 2 * px + c.pb;
      const bool valid = b < p.B && c.nkb > 0 && (p.rows_guard <= 0 || (long long)opix < p.rows_guard);
      const uint32_t tacc = tmem_base + buf * ACC + ((uint32_t)(q * 32) << 16);
      float vmin = INFINITY, vmax = -INFINITY;
      if (STAGED_OK && staged) {
        // bf16 outputs, 64 columns at a time: every lane packs its own pixel row (128 B) into the warp's swizzled
        // staging buffer, then 8 lanes write one row -- full 128-byte lines per store instead of 32 scattered
        // 16-byte pieces (the L1 tag stage processes one line per cycle, which bounded the thin layers).
        const uint32_t okm = __ballot_sync(0xffffffffu, valid);
        unsigned orow[8];                                 // (run_igemm refuses problems with 2^32 or more output pixels)
#pragma unroll
        for (int i = 0; i < 8; ++i) orow[i] = __shfl_sync(0xffffffffu, (unsigned)opix, 4 * i + (lane >> 3));
        // pad_out: this pixel's row index inside the bordered tensor, as an offset from opix (fetched by shuffle at the store)
        const unsigned dpad = p.pad_out ? (unsigned)((((size_t)b * (p.Hs + 2) + py + 1) * (p.Ws + 2) + px + 1) - opix) : 0u;
        const bool dpad_uniform = (p.Wt & 31) == 0;      // a warp's 32 pixels then lie in one image row: one offset for all of them
        float* const my_stats = ALT ? nullptr : stats_of(warp - 2);
        // padded: rows of the bordered tensor; cen: per-column constants subtracted after the activation (NULL: none);
        // col0 >= 0: accumulate the BatchNorm partial sums of the stored values for columns col0 .. col0 + 63
        auto emit = [&](const float (&v)[64], bf16* base, int ldn, float slope, bool act, bool padded, const float* cen, int col0) {
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
            unsigned row = orow[i];
            if (padded) row += dpad_uniform ? dpad : __shfl_sync(0xffffffffu, dpad, rr);
            if ((okm >> rr) & 1u) *reinterpret_cast<uint4*>(base + row * (unsigned long long)ldn + (lane & 7) * 8) = u;
          }
          if (col0 >= 0) {
            // BatchNorm partial sums of the stored (rounded) values: lane = (16-byte chunk j, word w) owns columns
            // 8j + 2w, 8j + 2w + 1 and walks the 32 staged rows (one conflict-free 4-byte read per row); the sums go to
            // this warp's private accumulators -- no shuffles, no atomics
            const int j = lane >> 2, w = lane & 3;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const uint32_t u = *reinterpret_cast<const uint32_t*>(stg + r * 128 + ((j ^ (r & 7)) << 4) + w * 4);
              const float lo = ((okm >> r) & 1u) ? __uint_as_float(u << 16) : 0.f;
              const float hi = ((okm >> r) & 1u) ? __uint_as_float(u & 0xffff0000u) : 0.f;
              s0 += lo; q0 = fmaf(lo, lo, q0);
              s1 += hi; q1 = fmaf(hi, hi, q1);
            }
            float2* ps = reinterpret_cast<float2*>(my_stats + col0 + 8 * j + 2 * w);
            float2* pq = reinterpret_cast<float2*>(my_stats + p.N + col0 + 8 * j + 2 * w);
            float2 a = *ps, b2 = *pq;
            a.x += s0; a.y += s1; b2.x += q0; b2.y += q1;
            *ps = a; *pq = b2;
          }
          __syncwarp();
        };
        float row_m = 0.f, row_il = 0.f, row_d = 0.f;
        if (ATT && p.epi >= 3 && !p.stat_by_col && valid) {
          if (p.epi == 3) { row_m = ordered_to_float(p.stat_m[opix]); row_il = 1.f / p.stat_l[opix]; }
          else row_d = p.delta[opix];
        }
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N; cc += 64) {
          float v[64];
          tmem_ld32(tacc + (uint32_t)cc, v);
          tmem_ld32(tacc + (uint32_t)cc + 32u, v + 32);
          const int n = c.n0 + cc;
          if (ATT && p.epi == 3) {
            if (p.stat_by_col) {
#pragma unroll
              for (int k4 = 0; k4 < 64; k4 += 4) {          // (full unroll: v[] must keep static indices to stay in registers)
                const int4 mm = *reinterpret_cast<const int4*>(p.stat_m + n + k4);
                const float4 ll = ld4(p.stat_l + n + k4);
                v[k4 + 0] = __expf(v[k4 + 0] * p.escale - ordered_to_float(mm.x)) * __frcp_rn(ll.x);
                v[k4 + 1] = __expf(v[k4 + 1] * p.escale - ordered_to_float(mm.y)) * __frcp_rn(ll.y);
                v[k4 + 2] = __expf(v[k4 + 2] * p.escale - ordered_to_float(mm.z)) * __frcp_rn(ll.z);
                v[k4 + 3] = __expf(v[k4 + 3] * p.escale - ordered_to_float(mm.w)) * __frcp_rn(ll.w);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 64; ++k) v[k] = __expf(v[k] * p.escale - row_m) * row_il;
            }
          } else if (ATT && p.epi == 4) {
            if (valid) {
              const bf16* prow = p.pmat + opix * (size_t)p.N + n;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float8 pv = ld8(prow + 8 * j);
                float dl[8];
                if (p.stat_by_col) {
                  const float4 d0 = ld4(p.delta + n + 8 * j), d1 = ld4(p.delta + n + 8 * j + 4);
                  dl[0] = d0.x; dl[1] = d0.y; dl[2] = d0.z; dl[3] = d0.w; dl[4] = d1.x; dl[5] = d1.y; dl[6] = d1.z; dl[7] = d1.w;
                } else {
#pragma unroll
                  for (int k = 0; k < 8; ++k) dl[k] = row_d;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) v[8 * j + k] = p.escale * pv.v[k] * (v[8 * j + k] - dl[k]);
              }
            }
          }
          if (p.act_dual) {
            if (p.aff_scale) {           // eval-mode BatchNorm: per-channel scale / shift of the running statistics
#pragma unroll
              for (int k4 = 0; k4 < 64; k4 += 4) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(p.aff_scale + n + k4));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(p.aff_shift + n + k4));
                v[k4] = fmaf(v[k4], sc.x, sh.x); v[k4 + 1] = fmaf(v[k4 + 1], sc.y, sh.y);
                v[k4 + 2] = fmaf(v[k4 + 2], sc.z, sh.z); v[k4 + 3] = fmaf(v[k4 + 3], sc.w, sh.w);
              }
            }
            emit(v, p.y0 + n, p.N, p.slope0, true, p.pad_out != 0, p.center ? p.center + n : nullptr, -1);
            if (p.y1) emit(v, p.y1 + n, p.N, p.slope1, true, false, nullptr, -1);
          } else if (n < p.N0) {
            emit(v, p.y0 + n, p.N0, 0.f, false, false, nullptr, want_stats ? n : -1);
          } else {
            emit(v, p.y1 + (n - p.N0), p.N1, 0.f, false, false, nullptr, want_stats ? n : -1);
          }
        }
      } else {
#pragma unroll 1
      for (int cc = egroup * 32; cc < BLOCK_N; cc += 32 * EG) {
        float v[32];
        tmem_ld32(tacc + (uint32_t)cc, v);
        if (EG == 2) {   // STFT instantiation only (mode 3)
          // row = frame px of waveform row b; columns (2k, 2k+1) = (Re, Im) of bin k -> |X| (optionally log)
          if (b < p.B && px < p.T) {
            const int k0 = (c.n0 + cc) >> 1;
            float* dst = p.spec + ((size_t)b * p.F + k0) * p.T + px;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (k0 + i < p.F) {
                float m = sqrtf(v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]);
                if (p.log_mode) {
                  m = __logf(m + 1e-8f);          // lg2.approx: absolute error ~1e-7 on a value range of ~10
                  vmin = fminf(vmin, m);
                  vmax = fmaxf(vmax, m);
                }
                dst[(size_t)i * p.T] = m;
              }
            }
          }
          continue;
        }
        if (!valid) continue;
        const int n = c.n0 + cc;
        if (ATT && p.epi == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) vmax = fmaxf(vmax, v[i] * p.escale);
        } else if (ATT && p.epi == 2) {
          if (cc == 0) { vmin = ordered_to_float(p.stat_m[opix]); vmax = 0.f; }    // (vmin = row max, vmax = running sum)
#pragma unroll
          for (int i = 0; i < 32; ++i) vmax += __expf(v[i] * p.escale - vmin);
        } else if (p.f32_rows && BLOCK_N != 16 && p.splits <= 1) {
          float* dst = p.out_f32 + opix * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else if (BLOCK_N == 16) {
          float* dst = p.out_f32 + opix * 16;
#pragma unroll
          for (int i = 0; i < 16; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else if (p.act_dual) {
          const size_t opad = p.pad_out ? ((size_t)b * (p.Hs + 2) + py + 1) * (p.Ws + 2) + px + 1 : opix;
          bf16* d0 = p.y0 + opad * p.N + n;
          bf16* d1 = p.y1 + opix * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = lrelu(v[i + k], p.slope0) - (p.center ? p.center[n + i + k] : 0.f);
            uint4 u, w;
            u.x = pack_bf16x2(a[0], a[1]); u.y = pack_bf16x2(a[2], a[3]);
            u.z = pack_bf16x2(a[4], a[5]); u.w = pack_bf16x2(a[6], a[7]);
            w.x = pack_bf16x2(lrelu(v[i + 0], p.slope1), lrelu(v[i + 1], p.slope1));
            w.y = pack_bf16x2(lrelu(v[i + 2], p.slope1), lrelu(v[i + 3], p.slope1));
            w.z = pack_bf16x2(lrelu(v[i + 4], p.slope1), lrelu(v[i + 5], p.slope1));
            w.w = pack_bf16x2(lrelu(v[i + 6], p.slope1), lrelu(v[i + 7], p.slope1));
            *reinterpret_cast<uint4*>(d0 + i) = u;
            *reinterpret_cast<uint4*>(d1 + i) = w;
          }
        } else if (p.splits > 1) {
          float* dst = p.partial + opix * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
        } else {
          bf16* dst = n < p.N0 ? p.y0 + opix * p.N0 + n : p.y1 + opix * p.N1 + (n - p.N0);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(v[i + 0], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
            u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
            *reinterpret_cast<uint4*>(dst + i) = u;
          }
        }
      }
      }
      if (ATT && EG == 1 && valid) {
        if (p.epi == 1) atomicMax(&p.stat_m[opix], float_to_ordered(vmax));
        else if (p.epi == 2) atomicAdd(&p.stat_l[opix], vmax);
      }
      if (EG == 2 && p.log_mode) {
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        if (lane == 0 && vmin <= vmax && b < p.B) {
          atomicMin(&p.minmax[2 * b], float_to_ordered(vmin));
          atomicMax(&p.minmax[2 * b + 1], float_to_ordered(vmax));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);      // this warp's quarter of the accumulator is free again
    }
  }
  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    // one fp64 atomic per (CTA, channel, moment); CTAs start at different channels so that they do not queue on one cell
    const int n2 = 2 * p.N;
    const int start = (int)(((unsigned)worker * 67u) % (unsigned)n2);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      int c = i + start;
      if (c >= n2) c -= n2;
      const float v = (stats_of(0)[c] + stats_of(1)[c]) + (stats_of(2)[c] + stats_of(3)[c]);
      if (v != 0.f) atomicAdd(&p.stats[c], (double)v);
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * ACC);
  }
}

// fp32 partial sums [pixels][N] -> bf16 outputs (split at N0); rezero: the sums are cleared as they are read, so a scratch
// that was zero before the convolution is zero again afterwards (the engine clears it once per step instead of per layer)
__global__ void __launch_bounds__(256)
finish_partial_kernel(float* __restrict__ partial, long long pixels, int N, int N0, int N1, bf16* __restrict__ y0,
                      bf16* __restrict__ y1, int rezero) {
  const long long n4 = pixels * N / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = 4 * i;
    const long long pix = e / N;
    const int n = (int)(e - pix * N);
    float4 v = ld4(partial + e);
    if (rezero) st4(partial + e, make_float4(0.f, 0.f, 0.f, 0.f));
    if (n < N0) st4(y0 + pix * N0 + n, v);
    else st4(y1 + pix * N1 + (n - N0), v);
  }
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int g_max_block_n = 256;   // "tc_max_bn" / ADP_TC_MAX_BN: largest N tile

int pick_block_n(int N, int N0, int N1) {
  const int cands[4] = {256, 128, 64, 32};
  for (int c : cands)
    if (c <= g_max_block_n && N % c == 0 && (N1 == 0 || N0 % c == 0)) return c;
  return 0;
}

// The deepest, weight-streaming layers (a handful of 128-pixel tiles, K split over the SMs): 128-wide N tiles halve what
// every CTA streams and reduces, and double the CTAs that share a K range (measured per launch at B = 64: E6 17.9 -> 14.7 us,
// E7 13.6 -> 10.4, D7 14.4 -> 11.5; from 64 tiles up -- E5, D6 -- the 256-wide tile wins again, as it does on every large layer)
int narrow_block_n_for_few_tiles(int bn, int m_tiles_x_par, int N, int N0, int N1) {
  if (bn != 256 || g_max_block_n < 256) return bn;
  if ((long long)m_tiles_x_par * (N / 256) > sm_count() / 4) return bn;
  if (N % 128 != 0 || (N1 != 0 && N0 % 128 != 0)) return bn;
  return 128;
}

bool tile_geometry(int B, int Hs, int Ws, int* Wt, int* Ht, int* Bt) {
  if (!pow2(Hs) || !pow2(Ws)) return false;
  *Wt = Ws < TILE_M ? Ws : TILE_M;
  int rest = TILE_M / *Wt;
  *Ht = Hs < rest ? Hs : rest;
  *Bt = rest / *Ht;
  (void)B;
  return *Bt <= 256;
}

int g_halo = 1;            // "tc_halo" / ADP_TC_HALO: 0 = one TMA box per tap, 1 = halo windows (narrow-N parity and 3x3 layers)
int g_stats = 1;           // "tc_stats" / ADP_TC_STATS: BatchNorm statistics accumulated by the convolution epilogue
int g_alt = 1;             // "tc_alt" / ADP_TC_ALT: two alternating epilogue groups for the thin pointwise layers
int g_skip_pad_taps = 1;   // "tc_skip_pad_taps" / ADP_TC_SKIP_PAD_TAPS: 1 x 1 bottleneck layers visit only the taps that see data

template <int BLOCK_N, int EG, bool ATT, int HALO, bool ALT = false>
int launch_persist(const IgemmParams& p, int ctas, cudaStream_t s) {
  using PS = PersistSmem<BLOCK_N, HALO, ALT>;
  if (p.stats && (p.N > PS::STATS_N || EG != 1 || BLOCK_N < 64)) {
    adp_set_error("tc igemm: fused statistics need 64 <= N <= %d here", PS::STATS_N);
    return ADP_ERR_ARG;
  }
  ADP_SMEM_ATTR((tc_igemm_persist_kernel<BLOCK_N, EG, ATT, HALO, ALT>), PS::BYTES);
  tc_igemm_persist_kernel<BLOCK_N, EG, ATT, HALO, ALT><<<ctas, 64 + 128 * EG, PS::BYTES, s>>>(p);
  return ADP_OK;
}

template <int BLOCK_N>
int launch_igemm(IgemmParams& p, dim3 grid, cudaStream_t s) {
  p.gx = grid.x; p.gy = grid.y; p.gz = grid.z;
  p.total_tiles = (int)(grid.x * grid.y * grid.z);
  const int sms = sm_count();
  const int ctas = p.total_tiles < sms ? p.total_tiles : sms;
  if (p.mode == 3) {
    if (BLOCK_N != 128) { adp_set_error("stft: BLOCK_N must be 128"); return ADP_ERR_ARG; }
    ADP_TRY((launch_persist<128, 2, false, 0>(p, ctas, s)));
  } else if (p.halo) {
    constexpr int HB = BLOCK_N == 128 ? 128 : 64;
    if (HB != BLOCK_N) { adp_set_error("halo mode needs BLOCK_N 64 or 128"); return ADP_ERR_ARG; }
    if (p.halo == 4) ADP_TRY((launch_persist<HB, 1, false, 4>(p, ctas, s)));
    else ADP_TRY((launch_persist<HB, 1, false, 3>(p, ctas, s)));
  } else if (p.epi != 0) {
    constexpr int AB = BLOCK_N < 64 ? 64 : BLOCK_N;
    if (AB != BLOCK_N) { adp_set_error("attention epilogues need BLOCK_N >= 64"); return ADP_ERR_ARG; }
    ADP_TRY((launch_persist<AB, 1, true, 0>(p, ctas, s)));
  } else if (g_alt && p.mode == 2 && (BLOCK_N == 64 || BLOCK_N == 128) && !p.f32_rows && !p.stats && p.splits <= 1 &&
             p.rows_guard <= 0 && p.kblocks <= 2 && (p.act_dual || p.N1 == 0 || p.N0 % 64 == 0)) {
    // thin pointwise layers (K <= 128): a 4-8 MMA main loop, the kernel is its epilogue -- two groups on alternate tiles
    constexpr int AB = (BLOCK_N == 64 || BLOCK_N == 128) ? BLOCK_N : 64;
    ADP_TRY((launch_persist<AB, 2, false, 0, true>(p, ctas, s)));
  } else {
    ADP_TRY((launch_persist<BLOCK_N, 1, false, 0>(p, ctas, s)));
  }
  adp_count_tc_launch();
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

// true when run_igemm would split the K range of this problem (fused statistics are then unavailable)
int pick_splits(const IgemmParams& p, int block_n, bool have_scratch, size_t scratch_bytes) {
  const int m_tiles = p.tiles_w * p.tiles_h * adp_cdiv(p.B, p.Bt);
  const int n_tiles = p.N / block_n;
  const int par = p.mode == 1 ? 4 : 1;
  const long long out_pixels = (long long)p.B * p.Hs * p.Ws * par;
  int splits = 1;
  const long long ctas = (long long)m_tiles * n_tiles * par;
  if ((have_scratch || p.f32_rows) && ctas < sm_count()) {
    splits = (int)(sm_count() / ctas);       // (rounded down: tiles x splits stays within one wave of the persistent grid)
    const int max_by_k = p.kblocks / 4 > 0 ? p.kblocks / 4 : 1;
    if (splits > max_by_k) splits = max_by_k;
    if (splits > 32) splits = 32;
    if (!p.f32_rows && (size_t)out_pixels * p.N * sizeof(float) > scratch_bytes) splits = 1;
    if (p.rows_guard > 0) splits = 1;      // (ragged row counts: keep the guarded direct store)
  }
  return splits;
}

int run_igemm(IgemmParams& p, int block_n, float* scratch, size_t scratch_bytes, cudaStream_t s, bool scratch_clean = false,
              float** deferred = nullptr) {
  const int m_tiles = p.tiles_w * p.tiles_h * adp_cdiv(p.B, p.Bt);
  const int n_tiles = p.N / block_n;
  const int par = p.mode == 1 ? 4 : 1;
  const long long out_pixels = (long long)p.B * p.Hs * p.Ws * par;
  ADP_CHECK_ARG(out_pixels + 4LL * p.B * (p.Hs + p.Ws + 4) < (1LL << 32), "tc igemm: too many output pixels (%lld)", out_pixels);
  // split the K range when the tile count cannot fill the machine (deep layers: small M, huge K)
  int splits = pick_splits(p, block_n, scratch != nullptr, scratch_bytes);
  p.kb_per_split = adp_cdiv(p.kblocks, splits);
  splits = adp_cdiv(p.kblocks, p.kb_per_split);
  p.splits = splits;
  if (splits > 1 && p.stats) { adp_set_error("tc igemm: fused statistics with a split K range"); return ADP_ERR_ARG; }
  if (p.f32_rows && splits > 1) { scratch = p.out_f32; scratch_clean = false; }   // fp32 result: the split-K partial sums ARE the output
  p.partial = splits > 1 ? scratch : nullptr;
  // (a clean scratch is all zero on entry and left all zero by finish_partial_kernel: no per-layer memset node)
  if (splits > 1 && !scratch_clean) ADP_CUDA(cudaMemsetAsync(scratch, 0, (size_t)out_pixels * p.N * sizeof(float), s));
  dim3 grid(m_tiles * par, n_tiles, splits);
  switch (block_n) {
    case 256: ADP_TRY(launch_igemm<256>(p, grid, s)); break;
    case 128: ADP_TRY(launch_igemm<128>(p, grid, s)); break;
    case 64: ADP_TRY(launch_igemm<64>(p, grid, s)); break;
    case 32: ADP_TRY(launch_igemm<32>(p, grid, s)); break;
    case 16: ADP_TRY(launch_igemm<16>(p, grid, s)); break;
    default: adp_set_error("tc igemm: bad BLOCK_N %d", block_n); return ADP_ERR_ARG;
  }
  if (splits > 1 && !p.f32_rows && scratch_clean && deferred) {
    *deferred = scratch;         // the consumer finishes (and clears) the sums
  } else if (splits > 1 && !p.f32_rows) {
    long long n4 = out_pixels * p.N / 4;
    int blocks = (int)((n4 + 255) / 256 < (long long)sm_count() * 8 ? (n4 + 255) / 256 : (long long)sm_count() * 8);
    finish_partial_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(scratch, out_pixels, p.N, p.N0, p.N1, p.y0, p.y1,
                                                                 scratch_clean ? 1 : 0);
    ADP_LAUNCH_CHECK();
  }
  return ADP_OK;
}

struct TcEnvInit {
  TcEnvInit() {
    const char* he = getenv("ADP_TC_HALO");
    if (he) g_halo = atoi(he);
    const char* be = getenv("ADP_TC_MAX_BN");
    if (be) g_max_block_n = atoi(be);
    const char* se = getenv("ADP_TC_STATS");
    if (se) g_stats = atoi(se);
    const char* ae = getenv("ADP_TC_ALT");
    if (ae) g_alt = atoi(ae);
    const char* pe = getenv("ADP_TC_SKIP_PAD_TAPS");
    if (pe) g_skip_pad_taps = atoi(pe);
  }
} g_tc_env_init;

// set and consumed within one C-ABI call on the calling host thread (thread_local: calls from several host threads,
// e.g. one per device, do not see each other's workspace)
thread_local float* g_scratch = nullptr;
thread_local size_t g_scratch_bytes = 0;
thread_local bool g_scratch_clean = false;      // the caller cleared it once and every split launch leaves it cleared

// eval-mode BatchNorm + activation(s) in the epilogue: needs the staged bf16 path of one un-split launch
bool fold_bn_act(IgemmParams& p, int block_n, const ConvExtras* ex) {
  if (!ex || !ex->bn_scale || !ex->bn_shift || !ex->y_act0 || block_n < 64 || p.N1 != 0) return false;
  if (pick_splits(p, block_n, g_scratch != nullptr, g_scratch_bytes) > 1) return false;
  p.act_dual = 1;
  p.aff_scale = ex->bn_scale; p.aff_shift = ex->bn_shift;
  p.slope0 = ex->slope0; p.slope1 = ex->slope1;
  p.y0 = reinterpret_cast<bf16*>(ex->y_act0);
  p.y1 = reinterpret_cast<bf16*>(ex->y_act1);
  p.N0 = p.N;
  if (ex->fold_done) *ex->fold_done = 1;
  return true;
}

// BatchNorm statistics can ride in the epilogue when the output goes through the staged bf16 path of one un-split launch
bool stats_fusable(const IgemmParams& p, int block_n, bool have_scratch, size_t scratch_bytes) {
  if (!g_stats || block_n < 64 || p.N1 != 0) return false;
  const int cap = p.halo == 3 ? 0 : ((p.halo && block_n == 64) ? PersistSmem<64, 4>::STATS_N : PersistSmem<256>::STATS_N);
  if (p.N > cap) return false;
  return pick_splits(p, block_n, have_scratch, scratch_bytes) <= 1;
}

}  // namespace

int tc_set_option(const char* name, int value) {
  int* slot = nullptr;
  if (!strcmp(name, "tc_halo")) slot = &g_halo;
  else if (!strcmp(name, "tc_stats")) slot = &g_stats;
  else if (!strcmp(name, "tc_alt")) slot = &g_alt;
  else if (!strcmp(name, "tc_skip_pad_taps")) slot = &g_skip_pad_taps;
  else if (!strcmp(name, "tc_max_bn")) slot = &g_max_block_n;
  if (!slot) return -1;
  const int prev = *slot;
  *slot = value;
  return prev;
}

void tc_set_scratch(void* ptr, size_t bytes, bool clean) {
  g_scratch = reinterpret_cast<float*>(ptr);
  g_scratch_bytes = bytes;
  g_scratch_clean = clean && ptr != nullptr;
}

bool tc_supported_gather(int B, int Hi, int Wi, int C, int N0, int N1) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !encode_tiled_fn()) return false;
  if (Hi % 2 || Wi % 2 || C % TILE_K || B < 1) return false;
  if (!tile_geometry(B, Hi / 2, Wi / 2, &Wt, &Ht, &Bt)) return false;
  return pick_block_n(N0 + N1, N0, N1) != 0;
}

bool tc_supported_parity(int B, int Hi, int Wi, int C0, int C1, int N) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !encode_tiled_fn()) return false;
  if (C0 % TILE_K || C1 % TILE_K || C0 <= 0 || B < 1) return false;
  if (!tile_geometry(B, Hi, Wi, &Wt, &Ht, &Bt)) return false;
  return pick_block_n(N, N, 0) >= 64;
}

int tc_gather_conv(const void* x, const void* w_nk, void* y0, int N0, void* y1, int N1, int B, int Hi, int Wi, int C,
                   cudaStream_t s, const ConvExtras* ex) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const int Ho = Hi / 2, Wo = Wi / 2, N = N0 + N1;
  ADP_CHECK_ARG(tile_geometry(B, Ho, Wo, &p.Wt, &p.Ht, &p.Bt), "tc_gather_conv: unsupported spatial size %dx%d", Hi, Wi);
  int bn = pick_block_n(N, N0, N1);
  ADP_CHECK_ARG(bn != 0 && C % TILE_K == 0, "tc_gather_conv: unsupported channels C=%d N0=%d N1=%d", C, N0, N1);
  p.tiles_w = Wo / p.Wt; p.tiles_h = Ho / p.Ht;
  if (g_scratch) bn = narrow_block_n_for_few_tiles(bn, p.tiles_w * p.tiles_h * adp_cdiv(B, p.Bt), N, N0, N1);
  p.B = B; p.Hs = Ho; p.Ws = Wo; p.mode = 0; p.C0 = C; p.C1 = 0; p.Ct = C; p.N = N; p.N0 = N0; p.N1 = N1;
  p.kblocks = 16 * (C / TILE_K);
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1;
  p.pad_in = ex && ex->pad_in ? 1 : 0;
  if (Ho == 1 && Wo == 1 && !p.pad_in && g_skip_pad_taps) {      // taps 5, 6, 9, 10 = (kh, kw) in {1, 2} x {1, 2}
    p.tap_lut = 0xA965u;
    p.kblocks = 4 * (C / TILE_K);
  }
  {  // x viewed as (2C | Wi/2 | 2 | Hi/2 | B); with an explicit border the tensor is [B, Hi+2, Wi+2, C]
    const int Hp = Hi + 2 * p.pad_in, Wp = Wi + 2 * p.pad_in;
    uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)Wp / 2, 2, (uint64_t)Hp / 2, (uint64_t)B};
    uint64_t str[4] = {(uint64_t)2 * C * 2, (uint64_t)Wp * C * 2, (uint64_t)2 * Wp * C * 2, (uint64_t)Hp * Wp * C * 2};
    uint32_t box[5] = {TILE_K, (uint32_t)p.Wt, 1, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(&p.tmA0, x, 5, dims, str, box));
  }
  {
    uint64_t dims[2] = {(uint64_t)16 * C, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)16 * C * 2};
    uint32_t box[2] = {TILE_K, (uint32_t)bn};
    ADP_TRY(make_tmap_bf16(&p.tmW, w_nk, 2, dims, str, box));
  }
  if (ex && ex->stats && stats_fusable(p, bn, g_scratch != nullptr, g_scratch_bytes)) {
    p.stats = ex->stats;
    if (ex->stats_done) *ex->stats_done = 1;
  }
  fold_bn_act(p, bn, ex);
  return run_igemm(p, bn, g_scratch, g_scratch_bytes, s, g_scratch_clean, ex ? ex->deferred : nullptr);
}

int tc_parity_convT(const void* x0, int C0, const void* x1, int C1, const void* w_kn, void* y, int B, int Hi, int Wi, int N,
                    cudaStream_t s, const ConvExtras* ex) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(tile_geometry(B, Hi, Wi, &p.Wt, &p.Ht, &p.Bt), "tc_parity_convT: unsupported spatial size %dx%d", Hi, Wi);
  int bn = pick_block_n(N, N, 0);
  ADP_CHECK_ARG(bn >= 64 && C0 % TILE_K == 0 && C1 % TILE_K == 0, "tc_parity_convT: unsupported channels");
  if (g_scratch) bn = narrow_block_n_for_few_tiles(bn, (Wi / p.Wt) * (Hi / p.Ht) * adp_cdiv(B, p.Bt) * 4, N, N, 0);
  const int Ct = C0 + C1;
  // narrow-N layers are bound by L2 -> SM operand traffic: load the tile's input window once per channel chunk (HALO)
  const bool halo = g_halo && (bn == 64 || bn == 128) && Hi >= 16 && Wi >= 8 && Hi % 16 == 0 && Wi % 8 == 0 &&
                    (long long)B * (Hi / 16) * (Wi / 8) * 4 * (N / bn) >= sm_count();
  if (halo) { p.Wt = 8; p.Ht = 16; p.Bt = 1; p.halo = 4; }
  p.tiles_w = Wi / p.Wt; p.tiles_h = Hi / p.Ht;
  p.B = B; p.Hs = Hi; p.Ws = Wi; p.mode = 1; p.C0 = C0; p.C1 = C1; p.Ct = Ct; p.N = N; p.N0 = N; p.N1 = 0;
  p.kblocks = (halo ? 1 : 4) * (Ct / TILE_K);
  if (Hi == 1 && Wi == 1 && !halo && g_skip_pad_taps) {
    p.one_tap = 1;
    p.kblocks = Ct / TILE_K;
  }
  p.y0 = (bf16*)y; p.y1 = nullptr;
  for (int h = 0; h < 2; ++h) {
    const int C = h == 0 ? C0 : C1;
    if (C == 0) continue;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Wi * C * 2, (uint64_t)Hi * Wi * C * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    uint32_t hbox[4] = {TILE_K, (uint32_t)HaloGeom<4>::W, HaloGeom<4>::H, 1};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : &p.tmA1, h == 0 ? x0 : x1, 4, dims, str, halo ? hbox : box));
  }
  {  // w_kn: bf16 [Ct][16][N] (the master layout, cast)
    uint64_t dims[3] = {(uint64_t)N, 16, (uint64_t)Ct};
    uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)16 * N * 2};
    uint32_t box[3] = {64, 1, TILE_K};
    ADP_TRY(make_tmap_bf16(&p.tmW, w_kn, 3, dims, str, box));
  }
  if (ex && ex->stats && stats_fusable(p, bn, g_scratch != nullptr, g_scratch_bytes)) {
    p.stats = ex->stats;
    if (ex->stats_done) *ex->stats_done = 1;
  }
  fold_bn_act(p, bn, ex);
  return run_igemm(p, bn, g_scratch, g_scratch_bytes, s, g_scratch_clean, ex ? ex->deferred : nullptr);
}

bool tc_supported_pointwise16(int B, int Hi, int Wi, int C0, int C1) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !encode_tiled_fn()) return false;
  if (C0 % TILE_K || C1 % TILE_K || C0 <= 0 || B < 1) return false;
  return tile_geometry(B, Hi, Wi, &Wt, &Ht, &Bt);
}

int tc_pointwise16(const void* x0, int C0, const void* x1, int C1, const void* w16, float* P, int B, int Hi, int Wi,
                   cudaStream_t s) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(tile_geometry(B, Hi, Wi, &p.Wt, &p.Ht, &p.Bt), "tc_pointwise16: unsupported spatial size %dx%d", Hi, Wi);
  ADP_CHECK_ARG(C0 % TILE_K == 0 && C1 % TILE_K == 0 && C0 > 0, "tc_pointwise16: unsupported channels");
  const int Ct = C0 + C1;
  p.tiles_w = Wi / p.Wt; p.tiles_h = Hi / p.Ht;
  p.B = B; p.Hs = Hi; p.Ws = Wi; p.mode = 2; p.C0 = C0; p.C1 = C1; p.Ct = Ct; p.N = 16; p.N0 = 16; p.N1 = 0;
  p.kblocks = Ct / TILE_K;
  p.out_f32 = P;
  for (int h = 0; h < 2; ++h) {
    const int C = h == 0 ? C0 : C1;
    if (C == 0) continue;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Wi * C * 2, (uint64_t)Hi * Wi * C * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : &p.tmA1, h == 0 ? x0 : x1, 4, dims, str, box));
  }
  {
    uint64_t dims[2] = {(uint64_t)Ct, 16};
    uint64_t str[1] = {(uint64_t)Ct * 2};
    uint32_t box[2] = {TILE_K, 16};
    ADP_TRY(make_tmap_bf16(&p.tmW, w16, 2, dims, str, box));
  }
  return run_igemm(p, 16, nullptr, 0, s);
}

// Pointwise (1x1) GEMM over NHWC pixels with K = C0 + C1 (multiple of 64):
//   y[pix][n] = sum_c (x0|x1)[pix][c] * w_nk[n][c]
// act_dual = 0: bf16 output split (y0: n < N0 | y1: rest);  act_dual = 1: y0 = lrelu(., slope0), y1 = lrelu(., slope1).
int tc_pointwise(const void* x0, int C0, const void* x1, int C1, const void* w_nk, void* y0, int N0, void* y1, int N1,
                 int act_dual, float slope0, float slope1, int B, int Hi, int Wi, cudaStream_t s, const ConvExtras* ex) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(tile_geometry(B, Hi, Wi, &p.Wt, &p.Ht, &p.Bt), "tc_pointwise: unsupported spatial size %dx%d", Hi, Wi);
  const int N = act_dual ? N0 : N0 + N1;
  const int bn = pick_block_n(N, N0, act_dual ? 0 : N1);
  ADP_CHECK_ARG(bn >= 32 && C0 % TILE_K == 0 && C1 % TILE_K == 0 && C0 > 0, "tc_pointwise: unsupported channels");
  const int Ct = C0 + C1;
  p.tiles_w = Wi / p.Wt; p.tiles_h = Hi / p.Ht;
  p.B = B; p.Hs = Hi; p.Ws = Wi; p.mode = 2; p.C0 = C0; p.C1 = C1; p.Ct = Ct; p.N = N; p.N0 = N0; p.N1 = act_dual ? 0 : N1;
  p.kblocks = Ct / TILE_K;
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1;
  p.act_dual = act_dual; p.slope0 = slope0; p.slope1 = slope1;
  if (ex && act_dual) { p.center = ex->center; p.pad_out = ex->pad_out ? 1 : 0; }
  for (int h = 0; h < 2; ++h) {
    const int C = h == 0 ? C0 : C1;
    if (C == 0) continue;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Wi * C * 2, (uint64_t)Hi * Wi * C * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : &p.tmA1, h == 0 ? x0 : x1, 4, dims, str, box));
  }
  {
    uint64_t dims[2] = {(uint64_t)Ct, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)Ct * 2};
    uint32_t box[2] = {TILE_K, (uint32_t)bn};
    ADP_TRY(make_tmap_bf16(&p.tmW, w_nk, 2, dims, str, box));
  }
  return run_igemm(p, bn, nullptr, 0, s);
}

// ------------------------------------------------------------------ 3x3 / stride 1 / pad 1 convolution (mode 4)
// models/binaural_attention_model.py:22-39 (DoubleConv).  Forward: w_nk = bf16 [N][9][Ct] (the nn.Conv2d weight in
// channels_last memory), wmode 0.  Data gradient: the SAME bf16 weight tensor [Cout][9][Cin] read through an
// MN-major (N = Cin | tap | K = Cout) map with the taps reversed (wmode 1), dy as the input, dx split N0 | N1.
bool tc_supported_conv3x3(int B, int H, int W, int C0, int C1, int N0, int N1) {
  int Wt, Ht, Bt;
  if (!adp_device_is_sm100() || !encode_tiled_fn()) return false;
  if (C0 <= 0 || C0 % TILE_K || C1 % TILE_K || B < 1) return false;
  if (!tile_geometry(B, H, W, &Wt, &Ht, &Bt)) return false;
  return pick_block_n(N0 + N1, N0, N1) >= 64;
}

int tc_conv3x3(const void* x0, int C0, const void* x1, int C1, const void* w, int wmode, void* y0, int N0, void* y1, int N1,
               int B, int H, int W, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(tile_geometry(B, H, W, &p.Wt, &p.Ht, &p.Bt), "tc_conv3x3: unsupported spatial size %dx%d", H, W);
  const int N = N0 + N1, Ct = C0 + C1;
  const int bn = pick_block_n(N, N0, N1);
  ADP_CHECK_ARG(bn >= 64 && C0 > 0 && C0 % TILE_K == 0 && C1 % TILE_K == 0, "tc_conv3x3: unsupported channels %d+%d -> %d+%d",
                C0, C1, N0, N1);
  // narrow-N layers: one (16 x 10)-pixel window per kernel row and channel chunk instead of one box per tap (HALO = 3)
  const bool halo = g_halo && (bn == 64 || bn == 128) && H >= 16 && W >= 8 && H % 16 == 0 && W % 8 == 0 &&
                    (long long)B * (H / 16) * (W / 8) * (N / bn) >= sm_count();
  if (halo) { p.Wt = 8; p.Ht = 16; p.Bt = 1; p.halo = 3; }
  p.tiles_w = W / p.Wt; p.tiles_h = H / p.Ht;
  p.B = B; p.Hs = H; p.Ws = W; p.mode = 4; p.wmode = wmode; p.C0 = C0; p.C1 = C1; p.Ct = Ct; p.N = N; p.N0 = N0; p.N1 = N1;
  p.kblocks = (halo ? 3 : 9) * (Ct / TILE_K);
  p.y0 = (bf16*)y0; p.y1 = (bf16*)y1;
  for (int h = 0; h < 2; ++h) {
    const int C = h == 0 ? C0 : C1;
    if (C == 0) continue;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Bt};
    uint32_t hbox[4] = {TILE_K, HaloGeom<3>::W, HaloGeom<3>::H, 1};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : &p.tmA1, h == 0 ? x0 : x1, 4, dims, str, halo ? hbox : box));
  }
  if (wmode) {   // w = [K = Ct][9][N]
    uint64_t dims[3] = {(uint64_t)N, 9, (uint64_t)Ct};
    uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)9 * N * 2};
    uint32_t box[3] = {64, 1, TILE_K};
    ADP_TRY(make_tmap_bf16(&p.tmW, w, 3, dims, str, box));
  } else {       // w = [N][9][Ct]
    uint64_t dims[2] = {(uint64_t)9 * Ct, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)9 * Ct * 2};
    uint32_t box[2] = {TILE_K, (uint32_t)bn};
    ADP_TRY(make_tmap_bf16(&p.tmW, w, 2, dims, str, box));
  }
  return run_igemm(p, bn, reinterpret_cast<float*>(scratch), scratch_bytes, s);
}

// ------------------------------------------------------------------ row GEMMs on the same kernel (mode 2)
//   C[m][n] = sum_k (A0|A1)[m][k] * Bop      m < M (any M >= 1), K = K0 + K1 (multiples of 64), N % 64 == 0
//   b_kn = 0: Bop = Bm[n][k] (K contiguous, "NT");  b_kn = 1: Bop = Bm[k][n] (N contiguous, "NN").
//   Output bf16 (c16, split N0 | N1 over two tensors) or fp32 (c32, one tensor).
// Used for the 1x1 convolutions and the attention products of models/binaural_attention_model.py:81-153.
int tc_gemm_rows(const void* a0, int K0, const void* a1, int K1, const void* bm, int b_kn, void* c16_0, int N0, void* c16_1,
                 int N1, float* c32, long long M, cudaStream_t s, const GemmEpilogue* epi) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  if (epi && epi->mode) {
    p.epi = epi->mode; p.stat_by_col = epi->by_col; p.escale = epi->scale;
    p.stat_m = epi->stat_m; p.stat_l = epi->stat_l; p.delta = epi->delta; p.pmat = reinterpret_cast<const bf16*>(epi->pmat);
  }
  ADP_CHECK_ARG(adp_device_is_sm100() && encode_tiled_fn(), "tc_gemm_rows: tcgen05 path unavailable");
  const int N = N0 + N1, Kt = K0 + K1;
  const int bn = pick_block_n(N, N0, N1);
  ADP_CHECK_ARG(bn >= 64 && K0 > 0 && K0 % TILE_K == 0 && K1 % TILE_K == 0 && M >= 1 && M < (1LL << 31),
                "tc_gemm_rows: unsupported shape M=%lld N=%d+%d K=%d+%d", M, N0, N1, K0, K1);
  if (p.epi == 1 || p.epi == 2) {
    ADP_CHECK_ARG(!c32 && !c16_0 && p.stat_m && (p.epi == 1 || p.stat_l) && bn >= 64, "tc_gemm_rows: reduction epilogue arguments");
  } else {
    ADP_CHECK_ARG((c32 != nullptr) != (c16_0 != nullptr), "tc_gemm_rows: exactly one of the bf16 / fp32 outputs");
    ADP_CHECK_ARG(p.epi == 0 || (c16_0 && N1 == 0 && (p.epi == 3 ? (p.stat_m && p.stat_l) : (p.delta && p.pmat))),
                  "tc_gemm_rows: softmax epilogue arguments");
  }
  const int mt = (int)((M + TILE_M - 1) / TILE_M);
  p.Wt = TILE_M; p.Ht = 1; p.Bt = 1; p.tiles_w = mt; p.tiles_h = 1;
  p.B = 1; p.Hs = 1; p.Ws = mt * TILE_M;           // one "image" of height 1 whose width is the (padded) row count
  p.mode = 2; p.wmode = b_kn; p.C0 = K0; p.C1 = K1; p.Ct = Kt; p.N = N; p.N0 = N0; p.N1 = N1;
  p.kblocks = Kt / TILE_K;
  p.rows_guard = M % TILE_M ? M : 0;
  p.y0 = (bf16*)c16_0; p.y1 = (bf16*)c16_1;
  p.out_f32 = c32; p.f32_rows = c32 ? 1 : 0;
  for (int h = 0; h < 2; ++h) {
    const int K = h == 0 ? K0 : K1;
    if (K == 0) continue;
    // rows beyond M are outside the tensor: zero-filled by TMA
    uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, 1, 1};
    uint64_t str[3] = {(uint64_t)K * 2, (uint64_t)M * K * 2, (uint64_t)M * K * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)TILE_M, 1, 1};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : &p.tmA1, h == 0 ? a0 : a1, 4, dims, str, box));
  }
  if (b_kn) {    // Bm = [K][N]
    uint64_t dims[3] = {(uint64_t)N, 1, (uint64_t)Kt};
    uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)N * 2};
    uint32_t box[3] = {64, 1, TILE_K};
    ADP_TRY(make_tmap_bf16(&p.tmW, bm, 3, dims, str, box));
  } else {       // Bm = [N][K]
    uint64_t dims[2] = {(uint64_t)Kt, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)Kt * 2};
    uint32_t box[2] = {TILE_K, (uint32_t)bn};
    ADP_TRY(make_tmap_bf16(&p.tmW, bm, 2, dims, str, box));
  }
  return run_igemm(p, bn, nullptr, 0, s);
}

// ------------------------------------------------------------------ STFT magnitude as a tensor-core DFT
// S[k,t] = | sum_{n<64} hann[n] x[hop*t + n - 32] e^{-2 pi i k n / n_fft} |  is the GEMM
//   [frames x 64] * [64 x 2F]  (columns = Re/Im of the windowed twiddles).  bf16 operands would cost 3 digits, so both
// operands are split x = hi + lo (two bf16 each) and the product is accumulated as hi*Whi + hi*Wlo + lo*Whi in fp32:
// Three-way splits (24 mantissa bits) and the six leading product terms: K = 384, fp32-level accuracy.
namespace {

__global__ void __launch_bounds__(256)
stft_frames_kernel(const float* __restrict__ wave, int L, int pitch, int hop, int T, int Tp, int rows,
                   bf16* __restrict__ A0, bf16* __restrict__ A1, bf16* __restrict__ A2) {
  const long long total = (long long)rows * Tp * 8;                 // one thread = 8 consecutive taps of one frame
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx & 7);
    const long long fr = idx >> 3;
    const int t = (int)(fr % Tp), row = (int)(fr / Tp);
    uint4 s0, s1, s2;
    uint32_t* p0 = &s0.x; uint32_t* p1 = &s1.x; uint32_t* p2 = &s2.x;
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float x[2], a[2], b[2], cc[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        x[u] = 0.f;
        if (t < T) {
          int j = hop * t + q * 8 + e + u - 32;
          if (j < 0) j = -j;
          if (j >= L) j = 2 * (L - 1) - j;
          j = min(max(j, 0), L - 1);
          x[u] = wave[(size_t)row * pitch + j];
        }
        a[u] = __bfloat162float(__float2bfloat16_rn(x[u]));
        b[u] = __bfloat162float(__float2bfloat16_rn(x[u] - a[u]));
        cc[u] = x[u] - a[u] - b[u];
      }
      p0[e / 2] = pack_bf16x2(a[0], a[1]);
      p1[e / 2] = pack_bf16x2(b[0], b[1]);
      p2[e / 2] = pack_bf16x2(cc[0], cc[1]);
    }
    *reinterpret_cast<uint4*>(A0 + idx * 8) = s0;
    *reinterpret_cast<uint4*>(A1 + idx * 8) = s1;
    *reinterpret_cast<uint4*>(A2 + idx * 8) = s2;
  }
}

// Wm bf16 [Npad][384]: row j = 2k (Re) / 2k+1 (Im); 64-column blocks W0, W1, W0, W2, W1, W0 (three-way split of w)
__global__ void stft_dft_matrix_kernel(bf16* __restrict__ Wm, int n_fft, int F, int Npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 64) return;
  const int j = i >> 6, n = i & 63;
  float w = 0.f;
  if (j < 2 * F) {
    const int k = j >> 1;
    const double hann = 0.5 - 0.5 * cospi(2.0 * n / 64.0);
    double sv, cv;
    sincospi(2.0 * (double)((k * n) % n_fft) / (double)n_fft, &sv, &cv);
    w = (float)(hann * ((j & 1) ? -sv : cv));
  }
  const float w0 = __bfloat162float(__float2bfloat16_rn(w));
  const float w1 = __bfloat162float(__float2bfloat16_rn(w - w0));
  const float w2 = w - w0 - w1;
  bf16* rowp = Wm + (size_t)j * 384;
  const bf16 b0 = __float2bfloat16_rn(w0), b1 = __float2bfloat16_rn(w1), b2 = __float2bfloat16_rn(w2);
  rowp[n] = b0; rowp[64 + n] = b1; rowp[128 + n] = b0; rowp[192 + n] = b2; rowp[256 + n] = b1; rowp[320 + n] = b0;
}

}  // namespace

size_t tc_stft_workspace_bytes(int rows, int L, int n_fft, int hop) {
  const int T = 1 + L / hop;
  int Tp = 128;
  while (Tp < T) Tp *= 2;
  const int Npad = ((2 * (n_fft / 2 + 1) + 127) / 128) * 128;
  return adp_align_up((size_t)rows * Tp * 64 * 2, 1024) * 3 + adp_align_up((size_t)Npad * 384 * 2, 1024);
}

bool tc_supported_stft(int rows, int L, int n_fft, int win, int hop) {
  if (!adp_device_is_sm100() || !encode_tiled_fn()) return false;
  return win == 64 && hop > 0 && n_fft >= 64 && n_fft <= 2048 && rows >= 1 && L > 32 && (1 + L / hop) <= 32768;
}

int tc_stft_mag(const float* wave, int rows, int L, int pitch, int n_fft, int hop, float* spec, int log_mode, int* minmax,
                void* workspace, cudaStream_t s) {
  const int T = 1 + L / hop, F = n_fft / 2 + 1;
  int Tp = 128;
  while (Tp < T) Tp *= 2;
  const int Npad = ((2 * F + 127) / 128) * 128;
  char* ws = reinterpret_cast<char*>(workspace);
  const size_t abytes = adp_align_up((size_t)rows * Tp * 64 * 2, 1024);
  bf16* A0 = reinterpret_cast<bf16*>(ws);
  bf16* A1 = reinterpret_cast<bf16*>(ws + abytes);
  bf16* A2 = reinterpret_cast<bf16*>(ws + 2 * abytes);
  bf16* Wm = reinterpret_cast<bf16*>(ws + 3 * abytes);
  {
    long long total = (long long)rows * Tp * 8, blocks = (total + 255) / 256, cap = (long long)sm_count() * 16;
    stft_frames_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, s>>>(wave, L, pitch, hop, T, Tp, rows, A0, A1, A2);
    ADP_LAUNCH_CHECK();
    stft_dft_matrix_kernel<<<adp_cdiv((long long)Npad * 64, 256), 256, 0, s>>>(Wm, n_fft, F, Npad);
    ADP_LAUNCH_CHECK();
  }
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  ADP_CHECK_ARG(tile_geometry(rows, 1, Tp, &p.Wt, &p.Ht, &p.Bt), "tc_stft: unsupported frame count %d", Tp);
  p.tiles_w = Tp / p.Wt; p.tiles_h = 1;
  p.B = rows; p.Hs = 1; p.Ws = Tp; p.mode = 3; p.C0 = 64; p.C1 = 64; p.Ct = 384; p.N = Npad; p.N0 = Npad; p.N1 = 0;
  p.kblocks = 6;
  p.spec = spec; p.F = F; p.T = T; p.log_mode = log_mode; p.minmax = minmax;
  for (int h = 0; h < 3; ++h) {
    uint64_t dims[4] = {64, (uint64_t)Tp, 1, (uint64_t)rows};
    uint64_t str[3] = {64 * 2, (uint64_t)Tp * 64 * 2, (uint64_t)Tp * 64 * 2};
    uint32_t box[4] = {TILE_K, (uint32_t)p.Wt, 1, 1};
    ADP_TRY(make_tmap_bf16(h == 0 ? &p.tmA0 : (h == 1 ? &p.tmA1 : &p.tmA2), h == 0 ? A0 : (h == 1 ? A1 : A2), 4, dims, str, box));
  }
  {
    uint64_t dims[2] = {384, (uint64_t)Npad};
    uint64_t str[1] = {384 * 2};
    uint32_t box[2] = {TILE_K, 128};
    ADP_TRY(make_tmap_bf16(&p.tmW, Wm, 2, dims, str, box));
  }
  return run_igemm(p, 128, nullptr, 0, s);
}

// wgrad on tensor cores: see adp_wgrad_tc.cu
}  // namespace adp
