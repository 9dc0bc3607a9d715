// p2p_common.cuh -- buffer layout and LL (data + epoch in one 8-byte store) primitives of the NVLink peer-memory
// exchanges (p2p_kernels.cu; the reduction fused into the CRD step's finalize kernel, crd_kernels.cu).
#pragma once
#include "common.cuh"

namespace crdpn {
namespace p2p {

constexpr int kMaxWorld = 8;
struct Peers { char* buf[kMaxWorld]; };
struct Offs { int off[kMaxWorld + 1]; };   // anchor offsets per rank (prefix sums of the per-rank batch sizes)

// layout of one rank's exchange buffer.  Payload areas hold 8-byte "LL" words {4 bytes of data, 4 bytes of epoch}:
// data and flag travel in ONE 8-byte store, so the receiver needs no fence and no separate flag -- it polls each
// word until its epoch matches (the protocol NCCL uses for small messages).
// Every payload area exists TWICE and an exchange of epoch e uses copy e & 1: a rank that runs ahead into the next
// exchange of the same kind writes the other copy, so it cannot overwrite words a slower peer has not read yet (it
// can only reach epoch e + 2 after that peer has finished epoch e, because its epoch e + 1 needs the peer's e + 1 push,
// which the peer issues -- in stream order -- after its epoch-e kernel).  Any call sequence is therefore safe, as long
// as it is the same on every rank.
struct Layout {
  size_t ctl, v1, v2, y, slots, parity_stride, slot_words, total;
  __host__ __device__ Layout(int64_t Bmax, int64_t Dmax, int world) {
    size_t o = 0;
    ctl = o; o += 64 * 4;        // [0] epoch of the gathers, [1] ticket, [2] epoch of the reductions, [3] ticket
    const size_t p0 = o;
    v1 = o; o += (size_t)Bmax * Dmax * 8;
    v2 = o; o += (size_t)Bmax * Dmax * 8;
    y = o; o += (size_t)Bmax * 2 * 8;
    o = (o + 255) / 256 * 256;
    // one reduction slot per rank: [grad_v1 (Bmax*Dmax) | grad_v2 (Bmax*Dmax) | 8 tail words | 8 scalars per anchor]
    slot_words = 2 * (size_t)Bmax * Dmax + 8 + 8 * (size_t)Bmax;
    slots = o; o += (size_t)world * slot_words * 8;
    o = (o + 255) / 256 * 256;
    parity_stride = o - p0;
    total = p0 + 2 * parity_stride;
  }
};

// how long a poll may last before the kernel gives up (SM clock ticks).  Ranks may legitimately drift apart by seconds
// (a checkpoint write, validation, a data-loader stall on one rank), so the default is ten minutes; a wait that long
// means a peer died or the call sequences diverged, and the context is then lost to a trap rather than to a hang.
// CRDPN_P2P_TIMEOUT_S (seconds, read once) overrides it.
long long poll_timeout_ticks();

__device__ __forceinline__ void ll_store(void* p, uint32_t data, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint32_t ll_load(const void* p, uint32_t epoch, long long timeout) {
  uint32_t d, f;
  const long long t0 = clock64();
  while (true) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(p) : "memory");
    if (f == epoch) break;
    if (clock64() - t0 > timeout) __trap();
  }
  return d;
}

}  // namespace p2p
}  // namespace crdpn
