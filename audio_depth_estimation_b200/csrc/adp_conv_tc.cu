// tcgen05 / TMEM / TMA implicit-GEMM convolutions (placeholder until the kernels land).
#include "adp_common.cuh"
namespace adp {
bool tc_supported_gather(int, int, int, int, int, int) { return false; }
bool tc_supported_parity(int, int, int, int, int, int) { return false; }
bool tc_supported_wgrad(int, int, int, int, int, int) { return false; }
int tc_gather_conv(const void*, const void*, void*, int, void*, int, int, int, int, int, cudaStream_t) {
  adp_set_error("tc_gather_conv: not built"); return ADP_ERR_UNSUPPORTED; }
int tc_parity_convT(const void*, int, const void*, int, const void*, void*, int, int, int, int, cudaStream_t) {
  adp_set_error("tc_parity_convT: not built"); return ADP_ERR_UNSUPPORTED; }
int tc_wgrad(const void*, int, const void*, int, const void*, int, float*, int, int, int, cudaStream_t) {
  adp_set_error("tc_wgrad: not built"); return ADP_ERR_UNSUPPORTED; }
}
