// p2p_kernels.cu -- the two exchanges of the row-sharded CRD step as single kernels over NVLink peer memory
// (one process per GPU, buffers shared with CUDA IPC), instead of two NCCL calls:
//   exchange 1  all-gather of the anchors' (v1, v2, idx) rows: every rank STORES its rows straight into every
//               peer's gather area and polls the peers' rows out of its own area -- one launch, one NVLink write
//               latency.
//   exchange 2  all-reduce of the packed partials [grad_v1 | grad_v2 | 8 scalars] (47 KB at B=46, D=128): one-shot
//               push of the partial into slot `rank` of every peer, then every rank sums the R slots in RANK ORDER
//               (deterministic, identical bits on every rank) -- one launch.
// Data and flag share one 8-byte store ("LL" words), so there is no fence and no separate flag round trip.
// Both payloads are far below the size where ring / tree algorithms pay off, so the cost is pure latency: the
// NCCL path costs two collective launches (~20-30 us each at 8 GPUs); these kernels cost one peer store + one flag.
// Epoch counters live in device memory and are advanced by the kernels themselves, so a captured CUDA graph can be
// replayed without patching arguments.  A wait that lasts ~2 s traps instead of hanging the box.
//
// Buffer reuse is safe without double buffering because the steps of a rank are stream-ordered and every exchange
// depends on data from all ranks: a peer can only push epoch e+1 after it has consumed this rank's epoch-e data.
#include <string.h>

#include "common.cuh"

namespace crdpn {
namespace p2p {

constexpr int kMaxWorld = 8;
struct Peers { char* buf[kMaxWorld]; };
struct Offs { int off[kMaxWorld + 1]; };   // anchor offsets per rank (prefix sums of the per-rank batch sizes)

// layout of one rank's exchange buffer.  Payload areas hold 8-byte "LL" words {4 bytes of data, 4 bytes of epoch}:
// data and flag travel in ONE 8-byte store, so the receiver needs no fence and no separate flag -- it polls each
// word until its epoch matches (the protocol NCCL uses for small messages).
struct Layout {
  size_t ctl, v1, v2, y, slots, total;
  __host__ __device__ Layout(int64_t Bmax, int64_t Dmax, int world) {
    size_t o = 0;
    ctl = o; o += 64 * 4;        // [0] epoch of the gathers, [1] ticket, [2] epoch of the reductions, [3] ticket
    v1 = o; o += (size_t)Bmax * Dmax * 8;
    v2 = o; o += (size_t)Bmax * Dmax * 8;
    y = o; o += (size_t)Bmax * 2 * 8;
    o = (o + 255) / 256 * 256;
    slots = o; o += (size_t)world * (2 * (size_t)Bmax * Dmax + 8) * 8;
    total = (o + 255) / 256 * 256;
  }
};

__device__ __forceinline__ void ll_store(void* p, uint32_t data, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint32_t ll_load(const void* p, uint32_t epoch) {
  uint32_t d, f;
  const long long t0 = clock64();
  while (true) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(p) : "memory");
    if (f == epoch) break;
    if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a peer died or the call sequences diverged
  }
  return d;
}
// push `nwords` 4-byte words to LL area `dst` (8 bytes per word) / poll them out of a local LL area
// part c of CH: the c-th of CH interleaved thread groups (blocks) covering the same word range
__device__ __forceinline__ void ll_push(char* dst, const void* src, size_t nwords, uint32_t e, int c, int CH) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (size_t i = (size_t)c * blockDim.x + threadIdx.x; i < nwords; i += (size_t)CH * blockDim.x) ll_store(dst + i * 8, s[i], e);
}
__device__ __forceinline__ void ll_pull(void* out, const char* src, size_t nwords, uint32_t e, int c, int CH) {
  uint32_t* d = reinterpret_cast<uint32_t*>(out);
  for (size_t i = (size_t)c * blockDim.x + threadIdx.x; i < nwords; i += (size_t)CH * blockDim.x) d[i] = ll_load(src + i * 8, e);
}
constexpr int kCH = 8;   // blocks per peer: the payloads are latency-bound, so spread the words over many threads

struct GatherParams {
  const float *v1, *v2;
  const long long* y;
  int D, rank, world;
  Offs offs;
  Peers peers;
  size_t off_ctl, off_v1, off_v2, off_y;
  float *out_v1, *out_v2;
  long long* out_y;
};

// grid = world x kCH blocks: blocks (p, *) push this rank's rows to peer p, then collect peer p's rows from this
// rank's own buffer
__global__ void __launch_bounds__(256) p2p_allgather_kernel(const GatherParams a) {
  const int p = blockIdx.x / kCH, c = blockIdx.x % kCH, tid = threadIdx.x;
  char* me = a.peers.buf[a.rank];
  uint32_t* ctl = reinterpret_cast<uint32_t*>(me + a.off_ctl);
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;   // advanced only after every block has read it
  const int D = a.D;
  {
    char* dst = a.peers.buf[p];
    const int a0 = a.offs.off[a.rank], n = a.offs.off[a.rank + 1] - a0;
    ll_push(dst + a.off_v1 + (size_t)a0 * D * 8, a.v1, (size_t)n * D, e, c, kCH);
    ll_push(dst + a.off_v2 + (size_t)a0 * D * 8, a.v2, (size_t)n * D, e, c, kCH);
    ll_push(dst + a.off_y + (size_t)a0 * 16, a.y, (size_t)n * 2, e, c, kCH);
  }
  {
    const int p0 = a.offs.off[p], n = a.offs.off[p + 1] - p0;
    ll_pull(a.out_v1 + (size_t)p0 * D, me + a.off_v1 + (size_t)p0 * D * 8, (size_t)n * D, e, c, kCH);
    ll_pull(a.out_v2 + (size_t)p0 * D, me + a.off_v2 + (size_t)p0 * D * 8, (size_t)n * D, e, c, kCH);
    ll_pull(a.out_y + p0, me + a.off_y + (size_t)p0 * 16, (size_t)n * 2, e, c, kCH);
  }
  __syncthreads();
  if (tid == 0) {  // the last block of this rank advances the epoch
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == (unsigned)(a.world * kCH - 1)) { ctl[1] = 0u; *reinterpret_cast<volatile uint32_t*>(ctl) = e; }
  }
}

struct ReduceParams {
  const float* partial;
  const double* tail;   // optional: n_tail doubles appended (as float) behind the n_main floats of `partial`
  float* out;
  int n, n_main, rank, world;
  Peers peers;
  size_t off_ctl, off_slots;
  size_t slot_stride;   // words between slots
};

// grid = world x kCH blocks: blocks (p, *) push the whole partial to peer p's slot `rank`; then block b reduces the
// b-th part of the payload over the R slots of its own buffer in rank order (every word is polled until it carries
// this epoch)
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const ReduceParams a) {
  const int p = blockIdx.x / kCH, c = blockIdx.x % kCH, tid = threadIdx.x;
  char* me = a.peers.buf[a.rank];
  uint32_t* ctl = reinterpret_cast<uint32_t*>(me + a.off_ctl) + 2;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
  {
    char* dst = a.peers.buf[p] + a.off_slots + (size_t)a.rank * a.slot_stride * 8;
    for (int i = c * blockDim.x + tid; i < a.n; i += kCH * blockDim.x) {
      const float v = i < a.n_main ? a.partial[i] : (float)a.tail[i - a.n_main];
      ll_store(dst + (size_t)i * 8, __float_as_uint(v), e);
    }
  }
  {
    const int nb = a.world * kCH, b = blockIdx.x;
    const int i0 = (int)((long long)a.n * b / nb), i1 = (int)((long long)a.n * (b + 1) / nb);
    const char* slots = me + a.off_slots;
    for (int i = i0 + tid; i < i1; i += blockDim.x) {
      float s = __uint_as_float(ll_load(slots + (size_t)i * 8, e));
      for (int r = 1; r < a.world; ++r)   // rank order: the same bits on every rank
        s += __uint_as_float(ll_load(slots + ((size_t)r * a.slot_stride + i) * 8, e));
      a.out[i] = s;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == (unsigned)(a.world * kCH - 1)) { ctl[1] = 0u; *reinterpret_cast<volatile uint32_t*>(ctl) = e; }
  }
}

}  // namespace p2p
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_p2p_buffer_bytes(int64_t Bmax, int64_t Dmax, int world, size_t* bytes) {
  if (!bytes || Bmax <= 0 || Dmax <= 0 || world < 1 || world > p2p::kMaxWorld)
    return fail(CRDPN_E_BADARG, "crdpn_p2p_buffer_bytes: bad argument (world <= 8)");
  *bytes = p2p::Layout(Bmax, Dmax, world).total;
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes == 0) return fail(CRDPN_E_BADARG, "crdpn_p2p_alloc: bad argument");
  CRDPN_CUDA(cudaMalloc(dev_ptr, bytes));
  CRDPN_CUDA(cudaMemset(*dev_ptr, 0, bytes));
  CRDPN_CUDA(cudaDeviceSynchronize());
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_free(void* dev_ptr) {
  if (dev_ptr) CRDPN_CUDA(cudaFree(dev_ptr));
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_export(void* dev_ptr, void* handle64_host) {
  if (!dev_ptr || !handle64_host) return fail(CRDPN_E_BADARG, "crdpn_p2p_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  CRDPN_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64_host), dev_ptr));
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_import(const void* handle64_host, void** peer_ptr) {
  if (!handle64_host || !peer_ptr) return fail(CRDPN_E_BADARG, "crdpn_p2p_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, sizeof(h));
  CRDPN_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_close(void* peer_ptr) {
  if (peer_ptr) CRDPN_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return CRDPN_OK;
}

static int fill_peers(void* const* peer_bufs_host, int rank, int world, p2p::Peers* out) {
  if (!peer_bufs_host || world < 1 || world > p2p::kMaxWorld || rank < 0 || rank >= world)
    return fail(CRDPN_E_BADARG, "crdpn_p2p: bad rank / world (world <= 8)");
  for (int r = 0; r < p2p::kMaxWorld; ++r) out->buf[r] = r < world ? (char*)peer_bufs_host[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!out->buf[r]) return fail(CRDPN_E_BADARG, "crdpn_p2p: null peer buffer");
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_allgather_anchors(const float* v1, const float* v2, const int64_t* y, int64_t D,
                                           const int32_t* offs_host, void* const* peer_bufs_host, int rank, int world,
                                           int64_t Bmax, int64_t Dmax, float* out_v1, float* out_v2, int64_t* out_y,
                                           void* stream) {
  if (!v1 || !v2 || !y || !offs_host || !out_v1 || !out_v2 || !out_y) return fail(CRDPN_E_BADARG, "crdpn_p2p_allgather_anchors: null pointer");
  p2p::GatherParams a;
  int rc = fill_peers(peer_bufs_host, rank, world, &a.peers);
  if (rc) return rc;
  if (D <= 0 || D > Dmax || offs_host[world] > Bmax || offs_host[0] != 0) return fail(CRDPN_E_BADARG, "crdpn_p2p_allgather_anchors: batch does not fit the exchange buffer");
  const p2p::Layout L(Bmax, Dmax, world);
  a.v1 = v1; a.v2 = v2; a.y = (const long long*)y; a.D = (int)D; a.rank = rank; a.world = world;
  for (int r = 0; r <= p2p::kMaxWorld; ++r) a.offs.off[r] = r <= world ? offs_host[r] : offs_host[world];
  a.off_ctl = L.ctl; a.off_v1 = L.v1; a.off_v2 = L.v2; a.off_y = L.y;
  a.out_v1 = out_v1; a.out_v2 = out_v2; a.out_y = (long long*)out_y;
  p2p::p2p_allgather_kernel<<<world * p2p::kCH, 256, 0, (cudaStream_t)stream>>>(a);
  CRDPN_LAUNCH_CHECK("p2p_allgather_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_allreduce_f32(const float* partial, int64_t n_main, const double* tail_f64, int64_t n_tail,
                                       float* out, void* const* peer_bufs_host,
                                       int rank, int world, int64_t Bmax, int64_t Dmax, void* stream) {
  if (!partial || !out || (n_tail > 0 && !tail_f64) || n_tail < 0) return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_f32: null pointer");
  const int64_t n = n_main + n_tail;
  p2p::ReduceParams a;
  int rc = fill_peers(peer_bufs_host, rank, world, &a.peers);
  if (rc) return rc;
  const size_t stride = 2 * (size_t)Bmax * Dmax + 8;
  if (n <= 0 || (size_t)n > stride) return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_f32: payload does not fit the exchange buffer");
  const p2p::Layout L(Bmax, Dmax, world);
  a.partial = partial; a.tail = tail_f64; a.out = out; a.n = (int)n; a.n_main = (int)n_main; a.rank = rank; a.world = world;
  a.off_ctl = L.ctl; a.off_slots = L.slots; a.slot_stride = stride;
  p2p::p2p_allreduce_kernel<<<world * p2p::kCH, 256, 0, (cudaStream_t)stream>>>(a);
  CRDPN_LAUNCH_CHECK("p2p_allreduce_kernel");
  return CRDPN_OK;
}
