// libadp_b200: error reporting, device queries and library-wide switches.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "adp_common.cuh"

static thread_local char g_err[1024] = "";

void adp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* adp_last_error(void) { return g_err; }
extern "C" int adp_version(void) { return 100; }

extern "C" int adp_device_is_sm100(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10 && minor == 0) ? 1 : 0;
}

static long long g_launches = 0;
void adp_count_launch() { ++g_launches; }
extern "C" long long adp_launch_count(void) { return g_launches; }
static long long g_tc_launches = 0;
void adp_count_tc_launch() { ++g_tc_launches; }
extern "C" long long adp_tc_launch_count(void) { return g_tc_launches; }

namespace adp {

// ---- optional event timing of kernel families ------------------------------------------
namespace {
constexpr int PROF_MAX = 32768;
bool g_prof_on = false;
int g_prof_n = 0;
cudaEvent_t g_prof_ev[PROF_MAX][2];
int g_prof_kind[PROF_MAX];
double g_prof_work[PROF_MAX];
int g_prof_created = 0;
}  // namespace

ProfScope::ProfScope(int kind, cudaStream_t s, double work) : slot(-1), stream(s) {
  if (!g_prof_on || g_prof_n >= PROF_MAX) return;
  if (g_prof_n >= g_prof_created) {
    if (cudaEventCreate(&g_prof_ev[g_prof_n][0]) != cudaSuccess || cudaEventCreate(&g_prof_ev[g_prof_n][1]) != cudaSuccess)
      return;
    g_prof_created = g_prof_n + 1;
  }
  slot = g_prof_n++;
  g_prof_kind[slot] = kind;
  g_prof_work[slot] = work;
  cudaEventRecord(g_prof_ev[slot][0], stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof_ev[slot][1], stream);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int g_tc = -1;
int tc_enabled() {
  if (g_tc < 0) {
    const char* e = getenv("ADP_TC");
    g_tc = (e && e[0] == '0') ? 0 : 1;
  }
  return g_tc;
}

}  // namespace adp

// Force the tensor-core (1) or SIMT (0) convolution path for bf16 tensors; returns the previous value.
extern "C" int adp_set_tensor_core(int on) {
  int prev = adp::tc_enabled();
  adp::g_tc = on ? 1 : 0;
  return prev;
}

// Tuning switches of the tensor-core kernels (the ADP_TC_* environment variables, at run time): "tc_halo" (halo-window
// parity kernels), "tc_cluster" (2-CTA weight multicast), "tc_max_bn" (largest N tile).  Returns the previous value.
extern "C" int adp_set_option(const char* name, int value) {
  if (!name) return -1;
  const int prev = adp::tc_set_option(name, value);
  return prev >= 0 ? prev : adp::unet_set_option(name, value);   // "side_stream": weight gradients on a side stream
}

// Start (on = 1, clears the record) or stop (on = 0) timing the convolution kernel families.
extern "C" int adp_profile_enable(int on) {
  adp::g_prof_on = on != 0;
  if (on) adp::g_prof_n = 0;
  return ADP_OK;
}
// After the stream has been synchronised: per family (ADP_PROF_* order: gather, parity, wgrad,
// thin first/last layers, elementwise) total milliseconds, total algorithmic work (FLOP, or bytes
// for elementwise) and number of timed calls.  Arrays have 5 entries.
extern "C" int adp_profile_read(double* ms, double* work, long long* calls) {
  ADP_CHECK_ARG(ms && work && calls, "profile_read: null pointer");
  for (int k = 0; k < adp::PROF_KINDS; ++k) { ms[k] = 0.0; work[k] = 0.0; calls[k] = 0; }
  for (int i = 0; i < adp::g_prof_n; ++i) {
    float t = 0.f;
    ADP_CUDA(cudaEventSynchronize(adp::g_prof_ev[i][1]));
    ADP_CUDA(cudaEventElapsedTime(&t, adp::g_prof_ev[i][0], adp::g_prof_ev[i][1]));
    ms[adp::g_prof_kind[i]] += t;
    work[adp::g_prof_kind[i]] += adp::g_prof_work[i];
    calls[adp::g_prof_kind[i]] += 1;
  }
  return ADP_OK;
}
