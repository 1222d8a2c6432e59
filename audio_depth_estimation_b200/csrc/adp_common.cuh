// Shared helpers for libadp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/adp_b200.h"

void adp_set_error(const char* fmt, ...);

#define ADP_CHECK_ARG(cond, ...)                         \
  do {                                                   \
    if (!(cond)) {                                       \
      adp_set_error(__VA_ARGS__);                        \
      return ADP_ERR_ARG;                                \
    }                                                    \
  } while (0)

#define ADP_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      adp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return ADP_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

void adp_count_launch();
void adp_count_tc_launch();

#define ADP_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    adp_count_launch();                                                             \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      adp_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return ADP_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define ADP_TRY(call)            \
  do {                           \
    int r_ = (call);             \
    if (r_ != ADP_OK) return r_; \
  } while (0)

static inline int adp_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t adp_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

typedef __nv_bfloat16 bf16;

// ---- 4-wide typed loads / stores (fp32: 16 B, bf16: 8 B) -------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
// ---- 8-wide typed loads / stores (fp32: 2 x 16 B, bf16: 16 B) ---------------
struct float8 { float v[8]; };
__device__ __forceinline__ float8 ld8(const float* p) {
  float8 r;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ float8 ld8(const bf16* p) {
  float8 r;
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st8(float* p, const float8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2_(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}
__device__ __forceinline__ void st8(bf16* p, const float8& r) {
  uint4 u;
  u.x = pack_bf16x2_(r.v[0], r.v[1]); u.y = pack_bf16x2_(r.v[2], r.v[3]);
  u.z = pack_bf16x2_(r.v[4], r.v[5]); u.w = pack_bf16x2_(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// raw (unconverted) 8-element loads: several can be put in flight at 4 registers each for bf16
struct raw8f { float4 a, b; };
__device__ __forceinline__ raw8f ldraw8(const float* p) {
  raw8f r;
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
  return r;
}
__device__ __forceinline__ uint4 ldraw8(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ float8 cvt8(const raw8f& r) {
  float8 o;
  o.v[0] = r.a.x; o.v[1] = r.a.y; o.v[2] = r.a.z; o.v[3] = r.a.w; o.v[4] = r.b.x; o.v[5] = r.b.y; o.v[6] = r.b.z; o.v[7] = r.b.w;
  return o;
}
__device__ __forceinline__ float8 cvt8(const uint4& u) {
  float8 o;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o.v[2 * i] = __uint_as_float(w[i] << 16);
    o.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return o;
}
template <class T> struct Raw8;
template <> struct Raw8<float> { typedef raw8f type; };
template <> struct Raw8<bf16> { typedef uint4 type; };
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}

__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.f ? z : z * slope; }
__device__ __forceinline__ float lrelu_grad(float z, float slope) { return z > 0.f ? 1.f : slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ordered-int encoding of floats for atomicMin/atomicMax
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// ---- internal cross-file entry points (host) -------------------------------
#include <atomic>
#define ADP_MAX_DEVICES 64

namespace adp {

int current_device();
int sm_count();
int ensure_smem_attr(const void* func, int bytes, std::atomic<unsigned long long>* done);
// opt a kernel into > 48 KB of dynamic shared memory once per device (not once per process)
#define ADP_SMEM_ATTR(func, bytes)                                                  \
  do {                                                                              \
    static std::atomic<unsigned long long> done_{0};                                \
    ADP_TRY(adp::ensure_smem_attr(reinterpret_cast<const void*>(func), (bytes), &done_)); \
  } while (0)

// Optional per-family device timing (CUDA events on the launching stream), read by bench.py.
// (PROF_*_DEEP: the small-M weight-streaming levels E5-E8 / D8-D6, reported next to the large layers that carry 90 % of the
// FLOPs; work = algorithmic FLOP for the convolution kinds, algorithmic bytes for PROF_THIN_BYTES-free PROF_ELEM)
enum ProfKind { PROF_GATHER = 0, PROF_PARITY, PROF_WGRAD, PROF_THIN, PROF_ELEM, PROF_GATHER_DEEP, PROF_PARITY_DEEP,
                PROF_WGRAD_DEEP, PROF_FEATURE, PROF_LOSS, PROF_OPTIM, PROF_KINDS };   // (the last three: work = bytes)
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(int kind, cudaStream_t s, double work);
  ~ProfScope();
};

// elementwise / BN (adp_elem.cu).  All tensors are [rows, C] (NHWC flattened), C % 4 == 0.
// sums[0:C] += sum x, sums[C:2C] += sum x^2
int bn_stats(int dtype, const void* x, long long rows, int C, double* sums, cudaStream_t s);
// training: batch statistics from sums (+ running-stat update); eval: running statistics.
struct BnFin {
  const double* sums;     // [2C] sum x, sum x^2 (training)
  double inv_rows;        // 1 / rows
  float unbias;           // rows / (rows - 1) (1 when rows == 1): biased -> unbiased variance for the running estimate
  const float *gamma, *beta;
  float *rm, *rv;         // running statistics (updated when training)
  int training;
  float eps, momentum;
  float *scale, *shift, *mean, *invstd;   // outputs, kept for the backward pass
  // x is stored as (true value - mean_offset[c]) (first-level centring, adp_unet.cu); NULL = no offset.  Batch statistics
  // of the stored tensor are used as they are; the running mean and the eval-mode shift refer to the true value.
  const float* mean_offset;
};
int bn_finalize(const BnFin& f, int C, cudaStream_t s);
// finalize + (z = x*scale+shift; out0 = lrelu(z, slope0); out1 (optional) = lrelu(z, slope1)) in one launch
int bn_affine_act(int dtype, const void* x, long long rows, int C, const BnFin& f, float slope0, void* out0, float slope1,
                  void* out1, cudaStream_t s);
// z = x*scale+shift (scale==NULL: z = x); out0 = lrelu(z, slope0); out1 (optional) = lrelu(z, slope1)
int affine_act(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
               float slope0, void* out0, float slope1, void* out1, cudaStream_t s);
// gz = gA*act0'(z) + gB*act1'(z) (either may be NULL); z = x*scale+shift (scale NULL: z = x).
// sums[0:C] += sum gz, sums[C:2C] += sum gz*xhat,  xhat = (x-mean)*invstd
int act_bn_bwd_reduce(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const void* gA, float slope0,
                      const void* gB, float slope1, double* sums, cudaStream_t s);
// mode 0: dx = gz (no BN); 1: dx = gz*scale (eval BN); 2: dx = scale*(gz - s1/M - xhat*s2/M) (batch stats)
// dgamma/dbeta (optional, both or none): dgamma[c] = sums[C+c], dbeta[c] = sums[c], written by the same launch
int act_bn_bwd_apply(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift,
                     const float* mean, const float* invstd, const void* gA, float slope0,
                     const void* gB, float slope1, const double* sums, int mode, void* dx, float* dgamma, float* dbeta,
                     cudaStream_t s);
// un-finished split-K sums (ConvExtras::deferred) of a layer without BatchNorm, n = rows * C elements, cleared on the way:
// mask == NULL: raw = round(sums), act = lrelu(raw, slope);  mask != NULL: act = round(sums) * lrelu'(mask, slope)
int finish_act(int dtype, float* partial, long long n, const void* mask, float slope, void* raw, void* act, cudaStream_t s);
// Small tensors (deep levels): the whole BatchNorm layer in ONE launch, each tensor read once (an 8-channel slab per block
// staged in shared memory).  tensors = 1 (forward) / 3 (backward) slabs must fit: bn_small_ok.
bool bn_small_ok(int dtype, long long rows, int C, int tensors);
// batch statistics (f.training must be 1) -> scale/shift/mean/invstd + running statistics -> out0 [, out1]
// partial != NULL: x has not been written yet -- partial holds the fp32 split-K sums [rows][C] a tensor-core convolution
// left un-finished (ConvExtras::deferred); they are rounded into x here and cleared
int bn_small_fwd(int dtype, const void* x, long long rows, int C, const BnFin& f, float slope0, void* out0, float slope1,
                 void* out1, cudaStream_t s, float* partial = nullptr);
// gA of bn_small_bwd taken from un-finished split-K sums: [rows][ld] fp32, this layer's C channels at column off; columns
// [0, side_c) (side_c <= off) are finished into side [rows][side_c] on the way (side may be NULL); all of it is cleared
struct BnSmallPartial {
  float* partial;
  int ld, off;
  void* side;
  int side_c;
};
// act_bn_bwd_reduce + act_bn_bwd_apply (mode 1 or 2); sums (optional, double [2C]) receives sum gz / sum gz*xhat
int bn_small_bwd(int dtype, const void* x, long long rows, int C, const float* scale, const float* shift, const float* mean,
                 const float* invstd, const void* gA, float slope0, const void* gB, float slope1, int mode, void* dx,
                 float* dgamma, float* dbeta, double* sums, cudaStream_t s, const BnSmallPartial* pp = nullptr);
// dgamma[c] = sums[C+c], dbeta[c] = sums[c]
int bn_param_grads(const double* sums, int C, float* dgamma, float* dbeta, cudaStream_t s);
// final head: y = act(u + bias); du = dy*act'(y) ; dbias[0] += sum du   (out_ch == 1)
int head_bwd(const float* y, const float* dy, long long n, int final_sigmoid, float* du, float* dbias,
             cudaStream_t s);
int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t s);
// src [R][16][C] fp32 -> dst [C][16][R] bf16
int cast_transpose_taps(const float* src, void* dst, int R, int C, cudaStream_t s);

// SIMT implicit-GEMM convolutions (adp_conv_simt.cu); T selected by dtype, fp32 accumulate,
// weights always fp32 in the master layouts.
// F1: y[b,oy,ox,n] = sum_{kh,kw,c} x[b,2oy-1+kh,2ox-1+kw,c] * w[n][kh*4+kw][c]; output split (y0: n<N0 | y1: rest)
int simt_gather_conv(int dtype, const void* x, const float* w, void* y0, int N0, void* y1, int N1,
                     int B, int Hi, int Wi, int C, cudaStream_t s);
// F2: y[b,2i+a,2j+bb,n] = sum_{2x2 taps,c} x[b,i+a-1+th,j+bb-1+tw,c] * w[c][kh*4+kw][n]; input concat (x0:C0 | x1:C1)
int simt_parity_convT(int dtype, const void* x0, int C0, const void* x1, int C1, const float* w, void* y,
                      int B, int Hi, int Wi, int N, cudaStream_t s);
// F3: dw[m][tap][n] += sum_{b,i,j} S[b,i,j,m] * G[b,2i-1+kh,2j-1+kw,n]; S = (s0:M0 | s1:M1) [B,Hs,Ws,*], G [B,2Hs,2Ws,N]
int simt_wgrad(int dtype, const void* s0, int M0, const void* s1, int M1, const void* g, int N,
               float* dw, int B, int Hs, int Ws, cudaStream_t s);
// first conv (tiny Cin): x NCHW fp32 [B,Cin,H,W], w [N][16][Cin] -> NHWC [B,H/2,W/2,N]:
// out0 = lrelu(conv, slope0), out1 (optional) = lrelu(conv, slope1)
int first_conv_fprop(int dtype, const float* x, const float* w, float slope0, void* out0, float slope1,
                     void* out1, int B, int H, int W, int Cin, int N, cudaStream_t s);
int first_conv_wgrad(int dtype, const float* x, const void* dy, float* dw, int B, int H, int W, int Cin, int N,
                     cudaStream_t s);
// last convT (Cout = 1): inputs (x0 | x1) NHWC [B,Hi,Wi,C0+C1], w [C][16] fp32.
// u = convT + bias; y = final act(u)  -> y fp32 [B,1,2Hi,2Wi]
int last_convT_fprop(int dtype, const void* x0, int C0, const void* x1, int C1, const float* w,
                     const float* bias, int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s);
// du fp32 [B,1,2Hi,2Wi] -> g0 [B,Hi,Wi,C0], g1 [B,Hi,Wi,C1]
int last_convT_dgrad(int dtype, const float* du, const float* w, void* g0, int C0, void* g1, int C1,
                     int B, int Hi, int Wi, cudaStream_t s);
int last_convT_wgrad(int dtype, const void* x0, int C0, const void* x1, int C1, const float* du,
                     float* dw, int B, int Hi, int Wi, cudaStream_t s);

// tiled thin-layer kernels (adp_thin.cu)
bool thin_first_supported(int Cin, int N);
int thin_first_conv_fprop(int dtype, const float* x, const float* w, float slope0, void* out0, float slope1, void* out1,
                          int B, int H, int W, int Cin, int N, cudaStream_t s);
int thin_first_conv_wgrad(int dtype, const float* x, const void* dy, float* dw, int B, int H, int W, int Cin, int N,
                          cudaStream_t s);
int thin_last_convT_dgrad(int dtype, const float* du, const float* w, void* g0, int C0, void* g1, int C1, int B, int Hi,
                          int Wi, cudaStream_t s);
int thin_last_convT_wgrad(int dtype, const void* x0, int C0, const void* x1, int C1, const float* du, float* dw, int B,
                          int Hi, int Wi, cudaStream_t s);
// tensor-core route for the thin layers: bf16 patch rows [pixels][64], padded weights, result folding
int thin_patch_rows(const float* img, void* out, int B, int Cin, int H, int W, int split, cudaStream_t s);
// first-level centring (adp_unet.cu: use_center).  xsum[ci] += sum of plane ci over the batch (x NCHW fp32)
int center_input_sums(const float* x, int B, int Cin, long long HW, double* xsum, cudaStream_t s);
// m[n] = bf16(LeakyReLU_0.2(sum_{tap,ci} w1[n][tap][ci] * mean_x[ci]))  (n < 64, w1 fp32 [64][16][Cin]);
// T[n2] = sum_{tap,c} w2b[n2][tap][c] * m[c]  (w2b bf16 [N2][16][64]);  border ring of a0pad bf16 [B, H+2, W+2, 64] = -m
int center_tables(const double* xsum, double inv_count, const float* w1, int Cin, const void* w2b, int N2, float* m,
                  float* T, void* a0pad, int B, int H, int W, cudaStream_t s);
// dw[n][tap][c] += m[c] * scale[n] * gsum[n]   (dw fp32 [N2][16][C])
int center_wgrad_fix(float* dw, const float* m, const float* scale, const double* gsum, int N2, int C, cudaStream_t s);
// thin layers with the patch tile built in shared memory (adp_thin_tc.cu): Cin = 2 (network input, hi/lo split) and
// Cin = 1 (dL/du of the head), power-of-two output grids >= 16 wide
bool thin_tc_supported(int B, int Cin, int H, int W);
int thin_tc_first_conv(const float* x, const void* w_pad, void* a, float slope0, void* r, float slope1, const float* center,
                       int pad_out, int B, int H, int W, cudaStream_t s);
int thin_tc_first_wgrad(const float* x, const void* g_e, float* dw, int B, int H, int W, cudaStream_t s);
int thin_tc_first_wgrad_act(const float* x, const void* gA, const void* gB, const void* r, float slope, float* dw, int B, int H,
                            int W, cudaStream_t s);
int thin_tc_last_dgrad(const float* du, const void* w_pad, void* g0, void* g1, int B, int Hi, int Wi, cudaStream_t s);
// x1_scale / x1_shift given: x1 is t, the raw output of the previous transposed conv, and q = ReLU(t * scale + shift) is
// formed in shared memory (the forward pass then never wrote q: thin_tc_last_fwd)
int thin_tc_last_wgrad(const void* x0, const void* x1, const float* du, float* dw, int B, int Hi, int Wi, cudaStream_t s,
                       const float* x1_scale = nullptr, const float* x1_shift = nullptr);
bool thin_tc_last_fwd_supported(int B, int Hi, int Wi);
int thin_tc_last_fwd(const void* x0, const void* x1, const float* scale, const float* shift, const void* w16, const float* bias,
                     int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s);
int thin_pad_rows(const float* src, void* dst, int R, int K, int dup, cudaStream_t s);
int thin_fold_wgrad(const float* D, float* dw, int mode, int K, cudaStream_t s);
// y[pix][n] = sum_c (x0|x1)[pix][c] * w_nk[n][c] over NHWC pixels (adp_conv_tc.cu)
// Optional extras of the tensor-core convolutions (all zero / NULL = none):
struct ConvExtras {
  double* stats;        // gather / parity: [2N] sum and sum of squares of the stored output per channel, ACCUMULATED by the
  int* stats_done;      //   epilogue when the launch qualifies (un-split, staged bf16 output); *stats_done = 1 then, else untouched
  int pad_in;           // gather: the input is [B, Hi+2, Wi+2, C] with an explicit one-pixel border
  int pad_out;          // pointwise act_dual: y0 is the interior of a [B, Hi+2, Wi+2, N] tensor
  const float* center;  // pointwise act_dual: y0 = lrelu(D, slope0) - center[n]
  // gather / parity, inference: eval-mode BatchNorm and the activation(s) in the epilogue -- instead of the raw output,
  // y_act0 = lrelu(D * bn_scale[n] + bn_shift[n], slope0) and (optional) y_act1 = lrelu(., slope1) are written, and
  // *fold_done = 1; a launch that does not qualify (split K range, narrow N) writes the raw output as usual
  const float* bn_scale; const float* bn_shift;
  float slope0, slope1;
  void* y_act0; void* y_act1;
  int* fold_done;
  // gather / parity, split-K launch on a scratch the engine keeps clean: leave the fp32 partial sums [pixels][N0+N1]
  // un-finished for the consumer (bn_small_fwd / bn_small_bwd / finish_act round and clear them: one launch less per layer);
  // *deferred = the sums then, else untouched
  float** deferred;
};
int tc_pointwise(const void* x0, int C0, const void* x1, int C1, const void* w_nk, void* y0, int N0, void* y1, int N1,
                 int act_dual, float slope0, float slope1, int B, int Hi, int Wi, cudaStream_t s,
                 const ConvExtras* ex = nullptr);
// D[128][NT] (fp32, zeroed by the callee) = sum_rows A[row][0:128] * Bm[row][0:NT]; A = two 64-column halves
// (a0: [rows][lda0] at column ca0, a1: [rows][lda1] at column ca1), Bm: [rows][ldb], all bf16 row-major (adp_wgrad_tc.cu)
int tc_gemm_tn(const void* a0, int lda0, int ca0, const void* a1, int lda1, int ca1, const void* bm, int ldb, int NT,
               long long rows, float* D, cudaStream_t s);
// dw[m][n] += sum_rows a[row][m] * bm[row][n]  (fp32 [M][ldd], caller zeroes; M, N multiples of 64): 1x1-conv weight gradients
int tc_gemm_tn_full(const void* a, int M, const void* bm, int N, float* dw, int ldd, long long rows, cudaStream_t s);
// P fp32 [B,Hi,Wi,16] -> y fp32 [B,1,2Hi,2Wi] = act(bias + col2im(P))
int last_convT_col2im(const float* P, const float* bias, int final_sigmoid, float* y, int B, int Hi, int Wi, cudaStream_t s);
// P[pixel][16 taps] = sum_c (x0|x1)[pixel][c] * w16[tap][c]   (w16: bf16 [16][C0+C1]) on tensor cores
bool tc_supported_pointwise16(int B, int Hi, int Wi, int C0, int C1);
int tc_pointwise16(const void* x0, int C0, const void* x1, int C1, const void* w16, float* P, int B, int Hi, int Wi,
                   cudaStream_t s);

// STFT magnitude as a split-bf16 DFT GEMM on tensor cores (win_length 64); spec [rows][F][T] fp32
bool tc_supported_stft(int rows, int L, int n_fft, int win, int hop);
size_t tc_stft_workspace_bytes(int rows, int L, int n_fft, int hop);
int tc_stft_mag(const float* wave, int rows, int L, int pitch, int n_fft, int hop, float* spec, int log_mode, int* minmax,
                void* workspace, cudaStream_t s);

// tcgen05 paths (adp_conv_tc.cu), bf16 operands, fp32 accumulate in TMEM.
// w_nk: bf16 [N][16][C] (K-major B operand)
int tc_gather_conv(const void* x, const void* w_nk, void* y0, int N0, void* y1, int N1,
                   int B, int Hi, int Wi, int C, cudaStream_t s, const ConvExtras* ex = nullptr);
// w_kn: bf16 [C0+C1][16][N] (N-major B operand: the master layout of the weight, cast)
int tc_parity_convT(const void* x0, int C0, const void* x1, int C1, const void* w_kn, void* y,
                    int B, int Hi, int Wi, int N, cudaStream_t s, const ConvExtras* ex = nullptr);
// g_pad = 1: G is [B, 2Hs+2, 2Ws+2, N] with an explicit one-pixel border (its values enter the sums)
int tc_wgrad(const void* s0, int M0, const void* s1, int M1, const void* g, int N,
             float* dw, int B, int Hs, int Ws, cudaStream_t s, int g_pad = 0);
// 3x3 / stride 1 / pad 1 on the same kernel (binaural_attention_model.py DoubleConv).  wmode 0: w = bf16 [N][9][Ct]
// (forward); wmode 1: w = bf16 [Ct][9][N] read MN-major with reversed taps (data gradient through the forward weight).
// scratch: fp32 [pixels][N] for split-K on small grids, or NULL.
bool tc_supported_conv3x3(int B, int H, int W, int C0, int C1, int N0, int N1);
int tc_conv3x3(const void* x0, int C0, const void* x1, int C1, const void* w, int wmode, void* y0, int N0, void* y1, int N1,
               int B, int H, int W, void* scratch, size_t scratch_bytes, cudaStream_t s);
bool tc_supported_wgrad3x3(int B, int H, int W, int M, int N);
int tc_wgrad3x3(const void* sgrad, int M, const void* g, int N, int ldn, int n_off, float* dw, int B, int H, int W,
                cudaStream_t s);
// C[m][n] = sum_k (A0|A1)[m][k] * (b_kn ? Bm[k][n] : Bm[n][k]); bf16 (split N0|N1) or fp32 output; any M >= 1
// optional attention epilogue (score tile D, row = m, column = n):  mode 1: stat_m[m] = max(., scale*D) (ordered int);
// 2: stat_l[m] += sum exp(scale*D - m[m]);  3: C = exp(scale*D - m[i]) / l[i];  4: C = scale * pmat[m][n] * (D - delta[i]);
// i = m, or n when by_col (the transposed score matrix).  Modes 1/2 write no matrix.
struct GemmEpilogue {
  int mode, by_col;
  float scale;
  int* stat_m; float* stat_l; const float* delta; const void* pmat;
};
int tc_gemm_rows(const void* a0, int K0, const void* a1, int K1, const void* bm, int b_kn, void* c16_0, int N0, void* c16_1,
                 int N1, float* c32, long long M, cudaStream_t s, const GemmEpilogue* epi = nullptr);
// fp32 scratch used to split the K range of deep, small-M layers across CTAs (NULL: never split)
// clean: the caller has zeroed it on the launching stream; split launches then skip their memset, and whoever finishes the
// sums (finish_partial_kernel or the deferred consumer) re-zeroes what was used
void tc_set_scratch(void* ptr, size_t bytes, bool clean = false);
// switches ("tc_halo", "tc_max_bn", "tc_stats"): returns the previous value, -1 for an unknown name
int tc_set_option(const char* name, int value);
int unet_set_option(const char* name, int value);      // "side_stream"
bool tc_supported_gather(int B, int Hi, int Wi, int C, int N0, int N1);
bool tc_supported_parity(int B, int Hi, int Wi, int C0, int C1, int N);
bool tc_supported_wgrad(int B, int Hs, int Ws, int M0, int M1, int N);
// 0 = SIMT only, 1 = tcgen05 where supported (default).  Read from ADP_TC env once, or set explicitly.
int tc_enabled();

}  // namespace adp
