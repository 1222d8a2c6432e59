// libadp_b200: error reporting, device queries and library-wide switches.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
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

// (relaxed atomics: several host threads, e.g. one per device, may launch concurrently)
static std::atomic<long long> g_launches{0};
void adp_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long adp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
static std::atomic<long long> g_tc_launches{0};
void adp_count_tc_launch() { g_tc_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long adp_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }

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

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev;
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
int sm_count() {
  static std::atomic<int> tab[ADP_MAX_DEVICES];
  const int dev = current_device() % ADP_MAX_DEVICES;
  int n = tab[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    tab[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device): `done` is the call site's bitmap of devices
int ensure_smem_attr(const void* func, int bytes, std::atomic<unsigned long long>* done) {
  const int dev = current_device() % ADP_MAX_DEVICES;
  if ((done->load(std::memory_order_relaxed) >> dev) & 1ull) return ADP_OK;
  ADP_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done->fetch_or(1ull << dev, std::memory_order_relaxed);
  return ADP_OK;
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

// Switches of the tensor-core kernels: "tc_halo" (halo-window parity / 3x3 kernels; 0 = one TMA box per tap, the
// A/B partner the parity tests compare against), "tc_max_bn" (largest N tile), "tc_stats" (BatchNorm statistics from
// the convolution epilogue).  Returns the previous value.
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
// After the stream has been synchronised: per family total milliseconds, total algorithmic work (FLOP for the
// convolution families, bytes for the BatchNorm / activation passes) and number of timed calls.  n entries, in the order
// {gather conv, parity convT, wgrad, thin first/last layers, BatchNorm + activation passes,
//  gather conv (deep levels), parity convT (deep levels), wgrad (deep levels), feature stage, loss, clip + AdamW}; the first
// three hold the LARGE layers when n > 5 and large + deep when n == 5 (the round-1 layout).
extern "C" int adp_profile_read_n(int n, double* ms, double* work, long long* calls) {
  ADP_CHECK_ARG(ms && work && calls && (n == 5 || n == adp::PROF_KINDS), "profile_read: null pointer or bad entry count");
  for (int k = 0; k < n; ++k) { ms[k] = 0.0; work[k] = 0.0; calls[k] = 0; }
  for (int i = 0; i < adp::g_prof_n; ++i) {
    float t = 0.f;
    ADP_CUDA(cudaEventSynchronize(adp::g_prof_ev[i][1]));
    ADP_CUDA(cudaEventElapsedTime(&t, adp::g_prof_ev[i][0], adp::g_prof_ev[i][1]));
    int k = adp::g_prof_kind[i];
    if (n == 5 && k >= adp::PROF_FEATURE) continue;
    if (n == 5 && k >= adp::PROF_GATHER_DEEP) k -= adp::PROF_GATHER_DEEP;
    ms[k] += t;
    work[k] += adp::g_prof_work[i];
    calls[k] += 1;
  }
  return ADP_OK;
}
extern "C" int adp_profile_read(double* ms, double* work, long long* calls) { return adp_profile_read_n(5, ms, work, calls); }
