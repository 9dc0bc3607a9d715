// common.cuh -- shared host/device helpers for libcrdpn_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/crdpn_b200.h"

namespace crdpn {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s (code %d)", what, code);
  return code;
}

inline int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

#define CRDPN_CUDA(call)                                            \
  do {                                                              \
    cudaError_t _e = (call);                                        \
    if (_e != cudaSuccess) return ::crdpn::cuda_fail(_e, #call);    \
  } while (0)

#define CRDPN_LAUNCH_CHECK(name)                                    \
  do {                                                              \
    ::crdpn::g_launches.fetch_add(1, std::memory_order_relaxed);    \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::crdpn::cuda_fail(_e, name);     \
  } while (0)

// optional per-launch event bracketing of the dominant kernels (see crdpn_timing_enable)
extern std::atomic<int> g_timing_on;
void timing_mark(int kernel_id, bool begin, cudaStream_t st);
struct ScopedKernelTimer {
  int id; cudaStream_t st; bool on;
  ScopedKernelTimer(int id_, cudaStream_t st_) : id(id_), st(st_), on(g_timing_on.load(std::memory_order_relaxed) != 0) {
    if (on) timing_mark(id, true, st);
  }
  ~ScopedKernelTimer() { if (on) timing_mark(id, false, st); }
};

struct DeviceInfo {
  int sms;
  int max_smem_optin;
};
// cached per device (host side)
int device_info(int device, DeviceInfo* out);

__device__ __forceinline__ float warp_sum_xor(float v, int from, int to_inclusive) {
  // xor-butterfly over lane-bit offsets from `from` down to `to_inclusive` (powers of two)
  for (int off = from; off >= to_inclusive; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Philox4x32-10 (Salmon et al., SC'11): 128-bit block `ctr` of the stream keyed by `seed`; the identical function is
// oracle/crd_oracle.c's philox4x32_10.
__device__ __forceinline__ void philox4x32_10(unsigned long long seed, unsigned long long ctr, unsigned out[4]) {
  unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0u, c3 = 0u;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

}  // namespace crdpn
