// p2p_kernels.cu -- the two exchanges of the row-sharded CRD step as single kernels over NVLink peer memory
// (one process per GPU, buffers shared with CUDA IPC), instead of two NCCL calls:
//   exchange 1  all-gather of the anchors' (v1, v2, idx) rows: every rank STORES its rows straight into every
//               peer's gather area and polls the peers' rows out of its own area -- one launch, one NVLink write
//               latency.
//   exchange 2  all-reduce of the packed partials [grad_v1 | grad_v2 | 8 scalars] (47 KB at B=46, D=128): one-shot
//               push of the partial into slot `rank` of every peer, then every rank sums the R slots in RANK ORDER
//               (deterministic, identical bits on every rank) -- one launch.  (crdpn_crd_step_sharded fuses this
//               exchange into the step's reduction kernel instead: crd_kernels.cu.)
// Data and flag share one 8-byte store ("LL" words), so there is no fence and no separate flag round trip.
// Both payloads are far below the size where ring / tree algorithms pay off, so the cost is pure latency: the
// NCCL path costs two collective launches (~20-30 us each at 8 GPUs); these kernels cost one peer store + one flag.
// Epoch counters live in device memory and are advanced by the kernels themselves, so a captured CUDA graph can be
// replayed without patching arguments.  Payload areas are double-buffered by epoch parity (p2p_common.cuh), so any call
// sequence is safe as long as every rank issues the same one; a poll that lasts longer than the (configurable,
// default ten minutes) time-out traps instead of hanging the box.
#include <stdlib.h>
#include <string.h>

#include "p2p_common.cuh"

namespace crdpn {
namespace p2p {

long long poll_timeout_ticks() {
  static const long long ticks = [] {
    const char* s = getenv("CRDPN_P2P_TIMEOUT_S");
    double sec = s ? atof(s) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    return (long long)(sec * 2.0e9);   // SM clock <= 2 GHz
  }();
  return ticks;
}

// push `nwords` 4-byte words to LL area `dst` (8 bytes per word) / poll them out of a local LL area
// part c of CH: the c-th of CH interleaved thread groups (blocks) covering the same word range
__device__ __forceinline__ void ll_push(char* dst, const void* src, size_t nwords, uint32_t e, int c, int CH) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (size_t i = (size_t)c * blockDim.x + threadIdx.x; i < nwords; i += (size_t)CH * blockDim.x) ll_store(dst + i * 8, s[i], e);
}
__device__ __forceinline__ void ll_pull(void* out, const char* src, size_t nwords, uint32_t e, int c, int CH, long long to) {
  uint32_t* d = reinterpret_cast<uint32_t*>(out);
  for (size_t i = (size_t)c * blockDim.x + threadIdx.x; i < nwords; i += (size_t)CH * blockDim.x) d[i] = ll_load(src + i * 8, e, to);
}
constexpr int kCH = 8;   // blocks per peer: the payloads are latency-bound, so spread the words over many threads

struct GatherParams {
  const float *v1, *v2;
  const long long* y;
  int D, rank, world;
  Offs offs;
  Peers peers;
  size_t off_ctl, off_v1, off_v2, off_y, parity_stride;
  long long timeout;
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
  const size_t par = (size_t)(e & 1u) * a.parity_stride;
  me += par;
  {
    char* dst = a.peers.buf[p] + par;
    const int a0 = a.offs.off[a.rank], n = a.offs.off[a.rank + 1] - a0;
    ll_push(dst + a.off_v1 + (size_t)a0 * D * 8, a.v1, (size_t)n * D, e, c, kCH);
    ll_push(dst + a.off_v2 + (size_t)a0 * D * 8, a.v2, (size_t)n * D, e, c, kCH);
    ll_push(dst + a.off_y + (size_t)a0 * 16, a.y, (size_t)n * 2, e, c, kCH);
  }
  {
    const int p0 = a.offs.off[p], n = a.offs.off[p + 1] - p0;
    ll_pull(a.out_v1 + (size_t)p0 * D, me + a.off_v1 + (size_t)p0 * D * 8, (size_t)n * D, e, c, kCH, a.timeout);
    ll_pull(a.out_v2 + (size_t)p0 * D, me + a.off_v2 + (size_t)p0 * D * 8, (size_t)n * D, e, c, kCH, a.timeout);
    ll_pull(a.out_y + p0, me + a.off_y + (size_t)p0 * 16, (size_t)n * 2, e, c, kCH, a.timeout);
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
  size_t off_ctl, off_slots, parity_stride;
  size_t slot_stride;   // words between slots
  long long timeout;
};

// grid = world x kCH blocks: blocks (p, *) push the whole partial to peer p's slot `rank`; then block b reduces the
// b-th part of the payload over the R slots of its own buffer in rank order (every word is polled until it carries
// this epoch)
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const ReduceParams a) {
  const int p = blockIdx.x / kCH, c = blockIdx.x % kCH, tid = threadIdx.x;
  char* me = a.peers.buf[a.rank];
  uint32_t* ctl = reinterpret_cast<uint32_t*>(me + a.off_ctl) + 2;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
  const size_t par = (size_t)(e & 1u) * a.parity_stride;
  {
    char* dst = a.peers.buf[p] + par + a.off_slots + (size_t)a.rank * a.slot_stride * 8;
    for (int i = c * blockDim.x + tid; i < a.n; i += kCH * blockDim.x) {
      const float v = i < a.n_main ? a.partial[i] : (float)a.tail[i - a.n_main];
      ll_store(dst + (size_t)i * 8, __float_as_uint(v), e);
    }
  }
  {
    const int nb = a.world * kCH, b = blockIdx.x;
    const int i0 = (int)((long long)a.n * b / nb), i1 = (int)((long long)a.n * (b + 1) / nb);
    const char* slots = me + par + a.off_slots;
    for (int i = i0 + tid; i < i1; i += blockDim.x) {
      float s = __uint_as_float(ll_load(slots + (size_t)i * 8, e, a.timeout));
      for (int r = 1; r < a.world; ++r)   // rank order: the same bits on every rank
        s += __uint_as_float(ll_load(slots + ((size_t)r * a.slot_stride + i) * 8, e, a.timeout));
      a.out[i] = s;
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == (unsigned)(a.world * kCH - 1)) { ctl[1] = 0u; *reinterpret_cast<volatile uint32_t*>(ctl) = e; }
  }
}


// In-place sum over the ranks of up to four blocks of float32 / float64 accumulators in ONE launch (the hand-offs of the
// rank-synchronised PointNet training step).  The payload travels as 32-bit LL words (a double as two); each word -- each
// PAIR of words of a double -- is owned by one thread, which reads it, pushes it into slot `rank` of every peer, polls the
// `world` slots of its own buffer and writes the rank-ordered sum back in place: no other thread touches that element, so
// the update needs no scratch copy, and every rank ends up with identical bits.
struct BlocksParams {
  char* ptr[4];
  int first[4];     // first 32-bit word of block i in the flattened payload
  int units[4];     // elements of block i
  int f64[4];
  int n_blocks, total_units;
  int rank, world;
  Peers peers;
  size_t off_ctl, off_slots, parity_stride, slot_stride;
  long long timeout;
};

__global__ void __launch_bounds__(256) p2p_allreduce_blocks_kernel(const BlocksParams a) {
  char* me = a.peers.buf[a.rank];
  uint32_t* ctl = reinterpret_cast<uint32_t*>(me + a.off_ctl) + 2;
  const uint32_t e = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
  const size_t par = (size_t)(e & 1u) * a.parity_stride + a.off_slots;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < a.total_units; u += gridDim.x * blockDim.x) {
    int b = 0, base_unit = 0;
    while (b + 1 < a.n_blocks && u >= base_unit + a.units[b]) { base_unit += a.units[b]; ++b; }
    const int i = u - base_unit;
    if (a.f64[b]) {
      double* p = reinterpret_cast<double*>(a.ptr[b]) + i;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(*p);
      const size_t w = (size_t)a.first[b] + 2 * (size_t)i;
      for (int r = 0; r < a.world; ++r) {
        char* dst = a.peers.buf[r] + par + ((size_t)a.rank * a.slot_stride + w) * 8;
        ll_store(dst, (uint32_t)bits, e);
        ll_store(dst + 8, (uint32_t)(bits >> 32), e);
      }
      double sum = 0.0;
      for (int r = 0; r < a.world; ++r) {   // rank order: the same bits on every rank
        const char* src = me + par + ((size_t)r * a.slot_stride + w) * 8;
        const unsigned long long lo = ll_load(src, e, a.timeout), hi = ll_load(src + 8, e, a.timeout);
        const double v = __longlong_as_double((long long)(lo | (hi << 32)));
        sum = r == 0 ? v : sum + v;
      }
      *p = sum;
    } else {
      float* p = reinterpret_cast<float*>(a.ptr[b]) + i;
      const uint32_t bits = __float_as_uint(*p);
      const size_t w = (size_t)a.first[b] + (size_t)i;
      for (int r = 0; r < a.world; ++r)
        ll_store(a.peers.buf[r] + par + ((size_t)a.rank * a.slot_stride + w) * 8, bits, e);
      float sum = 0.f;
      for (int r = 0; r < a.world; ++r) {
        const float v = __uint_as_float(ll_load(me + par + ((size_t)r * a.slot_stride + w) * 8, e, a.timeout));
        sum = r == 0 ? v : sum + v;
      }
      *p = sum;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctl + 1, 1u) == gridDim.x - 1) { ctl[1] = 0u; *reinterpret_cast<volatile uint32_t*>(ctl) = e; }
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
  a.off_ctl = L.ctl; a.off_v1 = L.v1; a.off_v2 = L.v2; a.off_y = L.y; a.parity_stride = L.parity_stride;
  a.timeout = p2p::poll_timeout_ticks();
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
  const p2p::Layout L(Bmax, Dmax, world);
  const size_t stride = L.slot_words;
  if (n <= 0 || (size_t)n > stride) return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_f32: payload does not fit the exchange buffer");
  a.partial = partial; a.tail = tail_f64; a.out = out; a.n = (int)n; a.n_main = (int)n_main; a.rank = rank; a.world = world;
  a.off_ctl = L.ctl; a.off_slots = L.slots; a.slot_stride = stride; a.parity_stride = L.parity_stride;
  a.timeout = p2p::poll_timeout_ticks();
  p2p::p2p_allreduce_kernel<<<world * p2p::kCH, 256, 0, (cudaStream_t)stream>>>(a);
  CRDPN_LAUNCH_CHECK("p2p_allreduce_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_p2p_allreduce_blocks(void* const* block_ptrs_host, const int64_t* counts_host, const int* is_f64_host,
                                          int n_blocks, void* const* peer_bufs_host, int rank, int world, int64_t Bmax,
                                          int64_t Dmax, void* stream) {
  if (!block_ptrs_host || !counts_host || !is_f64_host || n_blocks < 1 || n_blocks > 4)
    return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_blocks: 1..4 blocks");
  p2p::BlocksParams a;
  int rc = fill_peers(peer_bufs_host, rank, world, &a.peers);
  if (rc) return rc;
  const p2p::Layout L(Bmax, Dmax, world);
  size_t words = 0;
  int units = 0;
  for (int i = 0; i < 4; ++i) { a.ptr[i] = nullptr; a.first[i] = 0; a.units[i] = 0; a.f64[i] = 0; }
  for (int i = 0; i < n_blocks; ++i) {
    if (!block_ptrs_host[i] || counts_host[i] <= 0 || counts_host[i] >= (1 << 28)) return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_blocks: bad block");
    if ((uintptr_t)block_ptrs_host[i] & (is_f64_host[i] ? 7 : 3)) return fail(CRDPN_E_ALIGN, "crdpn_p2p_allreduce_blocks: block alignment");
    a.ptr[i] = (char*)block_ptrs_host[i];
    a.first[i] = (int)words;
    a.units[i] = (int)counts_host[i];
    a.f64[i] = is_f64_host[i] ? 1 : 0;
    words += (size_t)counts_host[i] * (is_f64_host[i] ? 2 : 1);
    units += (int)counts_host[i];
  }
  if (words > L.slot_words) return fail(CRDPN_E_BADARG, "crdpn_p2p_allreduce_blocks: payload does not fit the exchange buffer");
  a.n_blocks = n_blocks; a.total_units = units; a.rank = rank; a.world = world;
  a.off_ctl = L.ctl; a.off_slots = L.slots; a.parity_stride = L.parity_stride; a.slot_stride = L.slot_words;
  a.timeout = p2p::poll_timeout_ticks();
  int grid = (units + 255) / 256;
  if (grid > 148) grid = 148;
  p2p::p2p_allreduce_blocks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  CRDPN_LAUNCH_CHECK("p2p_allreduce_blocks_kernel");
  return CRDPN_OK;
}
