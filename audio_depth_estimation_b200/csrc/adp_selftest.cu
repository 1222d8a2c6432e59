// Hardware probe (not on any product path): can a SWIZZLE_128B K-major UMMA A operand start at a row that is not a
// multiple of 8 (a shifted window into a larger shared-memory tile)?  This is what a "load the input tile + halo once,
// run all filter taps from shared memory" convolution kernel needs (planned).
//   X: bf16 [160][64] rows, W: bf16 [64][64];  out[m][n] = sum_k X[m + shift][k] * W[n][k],  m < 128.
// The A descriptor starts at smem_A + shift*128 bytes; `base_offset` goes into descriptor bits [49,52).
#include "adp_tc.cuh"

namespace adp {
namespace {
using namespace tc;

constexpr int ST_ROWS = 160;

struct SelftestParams {
  CUtensorMap tmX, tmW;
  int shift, base_offset, sbo_bytes;
  float* out;
};

__global__ void __launch_bounds__(192, 1) umma_offset_probe_kernel(const __grid_constant__ SelftestParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_s = smem;                          // [160][128 B]
  unsigned char* b_s = smem + ST_ROWS * 128;          // [64][128 B]  (20480 is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&bars[0], ST_ROWS * 128 + 64 * 128);
      tma_load_2d(a_s, &p.tmX, &bars[0], 0, 0);
      tma_load_2d(b_s, &p.tmW, &bars[0], 0, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      uint64_t ad = umma_smem_desc(smem_u32(a_s) + (uint32_t)p.shift * 128u, 16, (uint32_t)p.sbo_bytes);
      ad |= (uint64_t)(p.base_offset & 7) << 49;
      const uint64_t bd = umma_smem_desc(smem_u32(b_s), 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k != 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    float v[32];
    for (int cc = 0; cc < 64; cc += 32) {
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      for (int i = 0; i < 32; ++i) p.out[(q * 32 + lane) * 64 + cc + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace
}  // namespace adp

extern "C" int adp_selftest_umma_offset(const void* x, const void* w, int shift, int base_offset, int sbo_bytes, float* out,
                                        void* stream) {
  using namespace adp;
  ADP_CHECK_ARG(x && w && out && shift >= 0 && shift + 128 <= ST_ROWS && sbo_bytes % 16 == 0, "selftest_umma_offset: bad arguments");
  SelftestParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {64, ST_ROWS};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, ST_ROWS};
    ADP_TRY(tc::make_tmap_bf16(&p.tmX, x, 2, dims, str, box));
  }
  {
    uint64_t dims[2] = {64, 64};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 64};
    ADP_TRY(tc::make_tmap_bf16(&p.tmW, w, 2, dims, str, box));
  }
  p.shift = shift; p.base_offset = base_offset; p.sbo_bytes = sbo_bytes; p.out = out;
  const int smem = ST_ROWS * 128 + 64 * 128 + 1024 + 64;
  umma_offset_probe_kernel<<<1, 192, smem, (cudaStream_t)stream>>>(p);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}
