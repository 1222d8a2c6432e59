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
