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

}  // namespace crdpn
