// host.cu -- host-side pieces of the C ABI: error/launch bookkeeping, device cache, alias-table build.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace crdpn {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

std::atomic<int> g_timing_on{0};
namespace {
struct EventPair { cudaEvent_t a = nullptr, b = nullptr; };
std::mutex g_tmu;
std::vector<EventPair> g_pairs[CRDPN_K_COUNT];
size_t g_used[CRDPN_K_COUNT] = {0};
constexpr size_t kMaxPairs = 1 << 15;
}  // namespace

void timing_mark(int kid, bool begin, cudaStream_t st) {
  if (kid < 0 || kid >= CRDPN_K_COUNT) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  auto& v = g_pairs[kid];
  size_t& used = g_used[kid];
  if (begin) {
    if (used >= kMaxPairs) return;
    if (used == v.size()) {
      EventPair ep;
      if (cudaEventCreate(&ep.a) != cudaSuccess || cudaEventCreate(&ep.b) != cudaSuccess) return;
      v.push_back(ep);
    }
    cudaEventRecord(v[used].a, st);
  } else {
    if (used >= v.size()) return;
    cudaEventRecord(v[used].b, st);
    ++used;
  }
}

int device_info(int device, DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  static bool have[64] = {false};
  if (device < 0) {
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  }
  if (device >= 64) return fail(CRDPN_E_BADARG, "device index out of range");
  std::lock_guard<std::mutex> lk(mu);
  if (!have[device]) {
    int sms = 0, smem = 0, major = 0;
    CRDPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CRDPN_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CRDPN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(CRDPN_E_UNSUPPORTED, "libcrdpn_b200 is built for sm_100a (B200) only");
    cache[device].sms = sms;
    cache[device].max_smem_optin = smem;
    have[device] = true;
  }
  *out = cache[device];
  return CRDPN_OK;
}

}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_abi_version(void) { return CRDPN_ABI_VERSION; }
extern "C" const char* crdpn_last_error(void) { return g_err; }
extern "C" uint64_t crdpn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int crdpn_timing_enable(int on) {
  g_timing_on.store(on ? 1 : 0);
  return CRDPN_OK;
}

extern "C" int crdpn_timing_read(int kid, double* total_ms, uint64_t* launches) {
  if (kid < 0 || kid >= CRDPN_K_COUNT || !total_ms || !launches) return fail(CRDPN_E_BADARG, "crdpn_timing_read: bad argument");
  std::lock_guard<std::mutex> lk(g_tmu);
  double tot = 0.0;
  for (size_t i = 0; i < g_used[kid]; ++i) {
    CRDPN_CUDA(cudaEventSynchronize(g_pairs[kid][i].b));
    float ms = 0.f;
    CRDPN_CUDA(cudaEventElapsedTime(&ms, g_pairs[kid][i].a, g_pairs[kid][i].b));
    tot += (double)ms;
  }
  *total_ms = tot;
  *launches = (uint64_t)g_used[kid];
  g_used[kid] = 0;
  return CRDPN_OK;
}

// Vose alias tables with the stack pairing of the published CRD sampler (AliasMethod.__init__): fp32
// arithmetic throughout; probabilities are normalised only when their sum exceeds 1.
extern "C" int crdpn_alias_build(const float* probs, int64_t n, float* prob, int64_t* alias) {
  if (!probs || !prob || !alias || n <= 0) return fail(CRDPN_E_BADARG, "crdpn_alias_build: bad argument");
  double total = 0.0;
  for (int64_t i = 0; i < n; ++i) total += (double)probs[i];
  const float totalf = (float)total;
  const bool norm = totalf > 1.0f;
  std::vector<int64_t> small_stack, large_stack;
  small_stack.reserve((size_t)n);
  large_stack.reserve((size_t)n);
  const float nf = (float)n;
  for (int64_t k = 0; k < n; ++k) {
    const float pk = norm ? probs[k] / totalf : probs[k];
    const float scaled = nf * pk;
    prob[k] = scaled;
    alias[k] = 0;
    (scaled < 1.0f ? small_stack : large_stack).push_back(k);
  }
  while (!small_stack.empty() && !large_stack.empty()) {
    const int64_t s = small_stack.back();
    small_stack.pop_back();
    const int64_t l = large_stack.back();
    large_stack.pop_back();
    alias[s] = l;
    const float rest = (prob[l] - 1.0f) + prob[s];
    prob[l] = rest;
    (rest < 1.0f ? small_stack : large_stack).push_back(l);
  }
  for (int64_t k : small_stack) prob[k] = 1.0f;
  for (int64_t k : large_stack) prob[k] = 1.0f;
  return CRDPN_OK;
}
