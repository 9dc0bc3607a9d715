// crd_kernels.cu -- CRD memory-bank NCE step on B200 (sm_100a).
//
// Work the published CRD algorithm does with index_select + bmm + exp + div + autograd + index_copy_
// (ContrastMemory.forward / ContrastLoss.forward; reference insertion point KD/common/base_class.py:387,
// KD/vision/vanilla/vanilla_kd.py:158-160) is done here in three launches:
//   1. crd_score_kernel     one pass over the B*(K+1) sampled bank rows: both dot products, exp, NCE loss
//                           terms and the closed-form dL/ds applied straight back onto the row while it is
//                           still in registers (grad_v accumulates in registers).  HBM-bound: each sampled
//                           row of each bank is read exactly once; nothing of size B*K*D is ever written.
//   2. crd_finalize_kernel  deterministic fixed-order reduction of the per-warp partials.
//   3. crd_update_kernel    momentum + L2 renormalisation of the B positive rows (after all scoring).
// crdpn_crd_step runs 2 and 3 as ONE launch (crd_finalize_update_kernel): two launches per training step, three with the
// band-sort pre-pass (crd_band_sort_kernel, variant | 0x400: the default for banks of about the L2's size and larger), which
// orders every anchor's list by row band so that all warps sweep the bank together and repeated rows are L2 hits.
//
// Data layout in HBM: a bank row is D contiguous elements (fp32 or bf16), 16-byte aligned, row pitch
// `row_stride` elements.  The Python module allocates both banks interleaved as [N][2][D] so that one
// sampled index touches one contiguous 2*D*sizeof(elem) span.
//
// Partitioning: the flat pair space P = B*K1 is cut into NW equal contiguous ranges, one per resident warp
// (grid = SMs x blocks/SM, all resident).  Warps never synchronise with each other.  A warp walks its range
// anchor by anchor; per anchor it scans 32 contrast indices at a time, drops entries that fall outside
// this rank's bank shard, compacts the survivors into a per-warp shared-memory queue and consumes the
// queue R rows x U steps at a time with all 2*CH*U 128-bit loads of a step group issued before first use.
#include <stdlib.h>
#include <string.h>
#include <cuda.h>   // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link dependency)
#include "common.cuh"
#include "p2p_common.cuh"        // LL exchange words: the all-reduce of the sharded step rides in the reduction kernel
#include "pointnet_common.cuh"   // tcgen05 / TMEM / mbarrier helpers (crd_tc_stream.cuh)

namespace crdpn {

constexpr int kWarps = 8;          // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kQueueCap = 128;     // per-warp compaction queue (entries, power of two)
constexpr int kSlotExtra = 8;      // scalar partials appended to each slot (ls, lt, se1, se2, cnt, pad)
constexpr int kMaxBlocksPerSM = 4; // upper bound used for workspace sizing

struct ScoreParams {
  const char* bank1;
  const char* bank2;
  long long row_stride_bytes;
  const float* v1;
  const float* v2;
  const long long* idx;
  // where the contrast indices come from: 0 = the int64 list `idx`; 1 = the int32 list `idx32` (half the bytes to copy and
  // scan); 2 = drawn HERE, entry (b, k) = draw_base + floor(u64(Philox block (seed, offset + b*K1 + k)) * draw_n / 2^64),
  // column 0 = y[b] -- bit for bit the list crdpn_alias_draw_contrast would write for uniform tables, without the 8 bytes
  // written and read back per entry and without the draw launch
  int idx_mode;
  const int* idx32;
  const long long* y;
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;   // optional device-resident addend to `offset` (CUDA-graph replays: see crd_loss.cu)
  long long draw_n, draw_base;
  int B, K1, D;
  long long row_begin, row_end;
  float k_exp;          // log2(e) / T
  float inv_Z1, inv_Z2; // 1/Z (full mode)
  float c;              // K*Pn + eps
  float inv_mPn;        // 1 / (K*Pn)
  float eps_over_mPn;   // (c - K*Pn) / (K*Pn)
  float inv_BT;         // 1 / (B*T)
  float* out_v1;
  float* out_v2;
  float* slots;
  int maxseg;
  unsigned int* ticket;
  long long nw;         // warps the pair space is cut over (<= resident warps; the rest idle)
  int cta_reduce;       // 1: the 8 warps of a CTA fold their partials in shared memory (fixed order) and the CTA writes
                        //    at most two slots (its range covers <= 2 anchors): 8x fewer partials for the reduction kernel
  // compact mode (row-sharded step): crd_shard_filter_kernel has already dropped the entries other shards own.  Unit
  // u = b * (NC + 1) + c of anchor b: c = 0 the positive (k = 0), c >= 1 the survivors of the kFilterChunk entries
  // k in [1 + (c-1) kFilterChunk, ...) in list order: local rows cl[(b NC + c-1) kFilterChunk ...], ucount[b NC + c-1] of them.
  // The warps cut the COMPACT space into equal ranges (exactly balanced, no scan), warp w writes slot w + b for anchor b.
  int compact;              // 0 off, 1 compact lists in list order, 3 band-sorted lists + interleaved blocks (see crd_band_sort_kernel)
  int prefetch;             // 1: consume() prefetches the next step's rows into the L2 (CRDPN_SCORE_PREFETCH=1; measured: -4..7 % on
                            // compact shards, +3 % on the band-sorted step whose rows are L2 hits already: off by default)
  int wpu;                  // compact == 3: warps per anchor
  unsigned band_mul;        // compact == 3: band of local row r = min(31, (r * band_mul) >> 32)
  int NC;
  const int* cl;
  const int* ucount;
  long long* anchor_start;   // [B + 1] out: start of every anchor in the compact space (for the reduction kernel)
};
constexpr int kFilterEPT = 16;        // entries per thread of the filter pre-pass (multiple of 4)
constexpr int kFilterChunk = 256 * kFilterEPT;   // entries per filter unit (one CTA of 256 threads)
constexpr int kMaxUnits = 4096;      // unit-count prefix lives in shared memory

struct FinalizeParams {
  const float* slots;
  int maxseg;
  long long NW;
  int group;            // warps per slot-writing unit: 1 (every warp writes its own slots) or kWarps (ScoreParams::cta_reduce)
  const long long* anchor_start;   // compact mode (ScoreParams::compact): anchor b owns [anchor_start[b], anchor_start[b+1]) of the
                                   // compact space cut over NW warps, warp w's partial for anchor b is slot w + b; else null
  int wpa;              // band-sorted mode (ScoreParams::compact == 3): warps per anchor, anchor b's partials are slots [b wpa, (b+1) wpa)
  int B, K1, D;
  int full;
  float* grad_v1;
  float* grad_v2;
  double* anchor_part;  // [B][8]
  double* result;       // [8]
  unsigned int* ticket;
  // row-sharded step (crdpn_crd_step_sharded): the sum over ranks rides in this kernel.  Block b pushes anchor b's
  // reduced gradient rows and loss partials into slot `rank` of every peer (LL words) and sums the `world` slots of its
  // own buffer in rank order into `reduced` [grad_v1 | grad_v2 | 8 result words]: no separate all-reduce launch.
  int xchg;
  int rank, world;
  p2p::Peers peers;
  size_t off_ctl, off_slots, parity_stride, slot_words;
  long long timeout;
  float* reduced;
};

// set by crdpn_crd_loss_forward{,_sharded} around a step whose sampler offset lives on the device (variant bit 0x4000)
thread_local const unsigned long long* g_sampler_offset_dev = nullptr;

int sharded_step_core(void* bank1, void* bank2, int64_t row_stride, int bank_dtype, void* const* peer_bufs_host, int rank,
                      int world, int64_t Bmax, int64_t Dmax, const int64_t* contrast_idx, int64_t B, int64_t K1, int64_t D,
                      int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end, float T, float Z1, float Z2,
                      float eps, float momentum, float one_minus_momentum, float* v1_all, float* v2_all, int64_t* y_all,
                      float* partial, double* result, float* reduced, void* workspace, size_t workspace_bytes, int variant,
                      void* stream, int idx_mode = 0, uint64_t seed = 0, uint64_t offset = 0, int64_t draw_n = 0,
                      int64_t draw_base = 0, const float* v1_local = nullptr, const float* v2_local = nullptr,
                      const int64_t* y_local = nullptr, const int32_t* offs_host = nullptr);

__device__ __forceinline__ uint4 ld16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log(1+x), x >= 0: alternating series below 2^-5 (truncation < 2e-9 relative), libm above.
__device__ __forceinline__ float log1p_pos(float x) {
  if (x < 0.03125f) {
    float t = fmaf(x, 0.2f, -0.25f);
    t = fmaf(x, t, 0.33333334f);
    t = fmaf(x, t, -0.5f);
    t = fmaf(x, t, 1.0f);
    return x * t;
  }
  return log1pf(x);
}

// packed pair arithmetic (sm_100 FFMA2: two fp32 FMAs per issue slot): acc.{lo,hi} += a.{lo,hi} * b.{lo,hi}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float lo2(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

// global row index of contrast entry `pos` = b * K1 + k (ScoreParams::idx_mode)
__device__ __forceinline__ long long contrast_entry(const ScoreParams& p, long long pos, int b, long long anchor_base) {
  if (p.idx_mode == 0) return p.idx[pos];
  if (p.idx_mode == 1) return (long long)p.idx32[pos];
  if (pos == anchor_base) return p.y[b];
  unsigned rr[4];
  philox4x32_10(p.seed, p.offset + (p.offset_dev != nullptr ? __ldg(p.offset_dev) : 0ull) + (unsigned long long)pos, rr);
  const unsigned long long bits = ((unsigned long long)rr[0] << 32) | (unsigned long long)rr[1];
  return p.draw_base + (long long)__umul64hi(bits, (unsigned long long)p.draw_n);
}

// Row-sharded step, pre-pass: one CTA per (anchor, chunk of kFilterChunk entries k >= 1).  Every thread loads its
// kFilterEPT entries at once (one memory latency for the whole chunk, the whole list in flight across the grid: B * NC CTAs
// are ONE wave at 5 CTAs per SM), the survivors -- entries whose row this shard owns -- are written in list order (ballots +
// a prefix over the (slice, warp) counts), so the scoring pass that follows neither scans nor skips anything and its
// result does not depend on timing.  Inside the scoring pass the same scan is a chain of dependent steps per warp that
// costs ~20 us of a 75 us kernel on a shard that owns 1/8 of the rows.
__global__ void __launch_bounds__(256, 5) crd_shard_filter_kernel(const ScoreParams p, int* __restrict__ cl, int* __restrict__ ucount) {
  constexpr int EPT = kFilterEPT, NCNT = EPT * 8, CPL = NCNT / 32;   // counts: one per (slice of 256 entries, warp)
  __shared__ int s_cnt[NCNT], s_pre[NCNT + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.NC, c = blockIdx.x - b * p.NC;
  const long long anchor_base = (long long)b * p.K1;
  const int k0 = 1 + c * kFilterChunk;
  int row[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int k = k0 + j * 256 + tid;
    long long r = -1;
    if (k < p.K1) r = contrast_entry(p, anchor_base + k, b, anchor_base);
    row[j] = (r >= p.row_begin && r < p.row_end) ? (int)(r - p.row_begin) : -1;
  }
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, row[j] >= 0);
    if (lane == 0) s_cnt[j * 8 + warp] = __popc(m);
  }
  __syncthreads();
  if (warp == 0) {   // exclusive prefix over the (slice, warp) counts, in list order: CPL consecutive counts per lane
    int cv[CPL], tot = 0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { cv[i] = s_cnt[lane * CPL + i]; tot += cv[i]; }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int run = inc - tot;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { s_pre[lane * CPL + i] = run; run += cv[i]; }
    if (lane == 31) s_pre[NCNT] = inc;
  }
  __syncthreads();
  int* dst = cl + (size_t)blockIdx.x * kFilterChunk;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, row[j] >= 0);
    if (row[j] >= 0) dst[s_pre[j * 8 + warp] + __popc(m & ((1u << lane) - 1u))] = row[j];
  }
  if (tid == 0) ucount[blockIdx.x] = s_pre[NCNT];
}

// Band-sorted compact lists (ScoreParams::compact == 3).  The gather kernel is at the HBM roofline of B*(K+1) row reads,
// but at the headline shape every resident row is drawn ~3x per step and the repeats come from HBM again (L2 hit rate 6 %):
// the warps walk 2 368 unrelated stretches of the lists at once, so the whole bank is the working set.  Here every
// (anchor, chunk) list is STABLY counting-sorted by row band (32 bands of the shard) and the scoring pass hands the warps of a
// unit interleaved 32-entry blocks of the sorted list: every warp then sweeps the bank from band 0 to band 31, all of them at
// the same pace, the working set is the band or two they are in, and a row's repeats within the step are L2 hits.
// Same shard filter as crd_shard_filter_kernel; order within a band = list order (deterministic: ballot ranks + a prefix
// over the (band, slice, warp) counts).
constexpr int kBands = 32;
// lanes of the warp whose 5-bit `bin` equals this lane's, among the lanes with `valid` set (five ballots: fixed cost, where
// match.any iterates over the distinct values -- nearly one per lane here)
__device__ __forceinline__ unsigned same_bin_mask(int bin, bool valid) {
  unsigned m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int bit = 0; bit < 5; ++bit) {
    const unsigned bal = __ballot_sync(0xffffffffu, (bin >> bit) & 1);
    m &= ((bin >> bit) & 1) ? bal : ~bal;
  }
  return m;
}
__device__ __forceinline__ int band_of(const ScoreParams& p, int row) {
  const int b = (int)(((unsigned long long)(unsigned)row * (unsigned long long)p.band_mul) >> 32);
  return b < kBands - 1 ? b : kBands - 1;
}
// SPARSE (a shard that owns a fraction of the rows): the survivors are first compacted in list order into shared memory
// (the filter kernel's ballots + prefix), and only ceil(survivors / 256) slices go through the sort -- 2 of 16 on an 8-way shard.
template <bool SPARSE>
__global__ void __launch_bounds__(256, SPARSE ? 4 : 5) crd_band_sort_kernel(const ScoreParams p, int* __restrict__ cl, int* __restrict__ ucount) {
  constexpr int EPT = kFilterEPT;
  __shared__ int s_cnt[kBands * EPT * 8];   // counts: [band][slice of 256 entries < NS][warp]
  __shared__ int s_wsum[8];
  __shared__ int s_surv[SPARSE ? kFilterChunk : 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.NC, c = blockIdx.x - b * p.NC;
  const long long anchor_base = (long long)b * p.K1;
  const int k0 = 1 + c * kFilterChunk;
  const unsigned below = (1u << lane) - 1u;
  int row[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int k = k0 + j * 256 + tid;
    long long r = -1;
    if (k < p.K1) r = contrast_entry(p, anchor_base + k, b, anchor_base);
    row[j] = (r >= p.row_begin && r < p.row_end) ? (int)(r - p.row_begin) : -1;
  }
  int NS = EPT;   // slices that take part in the sort (CTA-uniform)
  if constexpr (SPARSE) {
    // ---- survivors, compacted in list order (crd_shard_filter_kernel's scheme), then re-dealt 256 per slice
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const unsigned m = __ballot_sync(0xffffffffu, row[j] >= 0);
      if (lane == 0) s_cnt[j * 8 + warp] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
      constexpr int CPL = EPT * 8 / 32;
      int cv[CPL], tot = 0;
#pragma unroll
      for (int i = 0; i < CPL; ++i) { cv[i] = s_cnt[lane * CPL + i]; tot += cv[i]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      int run = inc - tot;
#pragma unroll
      for (int i = 0; i < CPL; ++i) { s_cnt[EPT * 8 + lane * CPL + i] = run; run += cv[i]; }
      if (lane == 31) s_wsum[0] = inc;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const unsigned m = __ballot_sync(0xffffffffu, row[j] >= 0);
      if (row[j] >= 0) s_surv[s_cnt[EPT * 8 + j * 8 + warp] + __popc(m & below)] = row[j];
    }
    __syncthreads();
    const int ns = s_wsum[0];
    NS = (ns + 255) >> 8;
#pragma unroll
    for (int j = 0; j < EPT; ++j) row[j] = (j < NS && j * 256 + tid < ns) ? s_surv[j * 256 + tid] : -1;
    __syncthreads();   // s_cnt / s_wsum are reused below
  }
  for (int i = tid; i < kBands * NS * 8; i += 256) s_cnt[i] = 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    if (j >= NS) break;
    const bool valid = row[j] >= 0;
    const int bin = valid ? band_of(p, row[j]) : 0;
    const unsigned m = same_bin_mask(bin, valid);
    if (valid && (m & below) == 0u) s_cnt[(bin * NS + j) * 8 + warp] = __popc(m);   // the group's first lane
  }
  __syncthreads();
  {   // exclusive prefix over the 256 * NS counts in (band, slice, warp) order: NS consecutive counts per thread
    int cv[EPT], tot = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) { cv[i] = i < NS ? s_cnt[tid * NS + i] : 0; tot += cv[i]; }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int run = inc - tot;
#pragma unroll
    for (int w = 0; w < 8; ++w) run += (w < warp) ? s_wsum[w] : 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (i < NS) s_cnt[tid * NS + i] = run;
      run += cv[i];
    }
    if (tid == 255) ucount[blockIdx.x] = run;
  }
  __syncthreads();
  int* dst = cl + (size_t)blockIdx.x * kFilterChunk;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    if (j >= NS) break;
    const bool valid = row[j] >= 0;
    const int bin = valid ? band_of(p, row[j]) : 0;
    const unsigned m = same_bin_mask(bin, valid);
    if (valid) dst[s_cnt[(bin * NS + j) * 8 + warp] + __popc(m & below)] = row[j];
  }
}

template <typename T> struct Unpack;
template <> struct Unpack<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ void run(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <> struct Unpack<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ void run(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
  }
};

// T: bank element; LPR: lanes cooperating on one row; CH: 16-byte chunks per lane per row;
// U: row-steps whose loads are issued together; BPS: resident CTAs per SM; FULL: loss+grad vs. sums only.
template <typename T, int LPR, int CH, int U, int BPS, bool FULL>
__global__ void __launch_bounds__(kThreads, BPS) crd_score_kernel(const ScoreParams p) {
  constexpr int VEC = Unpack<T>::VEC;
  constexpr int R = 32 / LPR;
  constexpr int NV = VEC * CH;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr bool kFast = FULL && (LPR == 16 || LPR == 32) && U == 4;   // transposed score reduction in consume()
  constexpr int kDupShift = (LPR == 32) ? 2 : 1;                       // ... lanes sharing one reduced value: 1 << kDupShift
  static_assert(R * U + 64 <= kQueueCap, "queue too small");

  __shared__ int2 queue_smem[kWarps][kQueueCap];
  constexpr int kD = NV * LPR;                      // feature dimension of this instantiation
  constexpr int kSW = 2 * kD + kSlotExtra;          // floats per slot
  constexpr bool kCanReduce = kD <= 256;            // [8 warps][2 anchors][slot] must fit static shared memory
  // one buffer, two mutually exclusive uses: the CTA-level fold of the warp partials, or compact mode's unit-count prefix
  constexpr int kRedBytes = kCanReduce ? kWarps * 2 * kSW * 4 : 16;
  constexpr int kPreBytes = (kMaxUnits + 1) * 4;
  __shared__ __align__(16) unsigned char shbuf[kRedBytes > kPreBytes ? kRedBytes : kPreBytes];
  float* red_smem = reinterpret_cast<float*>(shbuf);
  int* s_upre = reinterpret_cast<int*>(shbuf);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPR, j = lane % LPR;
  const long long NW = p.nw;
  const long long gw = (long long)blockIdx.x * kWarps + warp;
  const long long P = (long long)p.B * p.K1;
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.ticket = 0u;  // finalize runs after us in stream order
  const bool cta_red = kCanReduce && p.cta_reduce != 0;
  int b_first_cta = 0;
  if (cta_red) {
    for (int i = threadIdx.x; i < kWarps * 2 * kSW; i += kThreads) red_smem[i] = 0.f;
    const long long w0 = (long long)blockIdx.x * kWarps;
    b_first_cta = (int)((P * (w0 < NW ? w0 : NW) / NW) / p.K1);
    __syncthreads();
  } else if (gw >= NW && p.compact != 1) {
    return;
  }
  long long lo = gw < NW ? P * gw / NW : 0;
  long long hi = gw < NW ? P * (gw + 1) / NW : 0;
  int2* q = queue_smem[warp];
  const bool store_out = (p.out_v1 != nullptr);
  int seg = 0;

  // ---- compact mode: exclusive prefix of the unit counts (shared memory), then equal ranges of the COMPACT space ----
  __shared__ int s_wsum[kWarps];
  const bool compact = p.compact == 1;
  const bool banded = p.compact == 3;   // band-sorted lists, warp gw = (anchor gw / wpu, interleave slot gw % wpu)
  const int UPA = p.NC + 1;                 // units per anchor: the positive, then NC filtered chunks
  int cu = 0;                               // current unit
  if (compact) {   // (cta_reduce is off in this mode, so every thread is still here)
    const int NUt = p.B * UPA;
    constexpr int kPer = kMaxUnits / kThreads;   // 16 consecutive units per thread
    int cntv[kPer], local = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int u = threadIdx.x * kPer + i;
      int cv = 0;
      if (u < NUt) {
        const int ub = u / UPA, uc = u - ub * UPA;
        if (uc == 0) {
          const long long r = contrast_entry(p, (long long)ub * p.K1, ub, (long long)ub * p.K1);
          cv = (r >= p.row_begin && r < p.row_end) ? 1 : 0;
        } else {
          cv = p.ucount[ub * p.NC + uc - 1];
        }
      }
      cntv[i] = cv;
      local += cv;
    }
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wbase += (w < warp) ? s_wsum[w] : 0;
    int run = wbase + inc - local;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int u = threadIdx.x * kPer + i;
      if (u <= NUt) s_upre[u] = run;
      run += cntv[i];
    }
    __syncthreads();
    const long long Tc = s_upre[NUt];
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i <= p.B; i += kThreads) p.anchor_start[i] = s_upre[i * UPA];
    if (gw >= NW) return;
    lo = Tc * gw / NW;
    hi = Tc * (gw + 1) / NW;
    if (lo < hi) {   // largest unit whose start is <= lo (binary search; empty units are skipped in the loop below)
      int a = 0, z = NUt;
      while (z - a > 1) {
        const int mid = (a + z) >> 1;
        if ((long long)s_upre[mid] <= lo) a = mid; else z = mid;
      }
      cu = a;
    }
  }

  if (banded) { lo = 0; hi = 1; }   // one pass of the loop below: this warp's share of ONE unit
  while (lo < hi) {
    if (compact)
      while ((long long)s_upre[cu + 1] <= lo) ++cu;
    const int b = banded ? (int)(gw / p.wpu) : compact ? cu / UPA : (int)(lo / p.K1);
    const long long anchor_base = (long long)b * p.K1;
    const long long seg_hi = banded ? 1
                           : compact ? ((hi < (long long)s_upre[(b + 1) * UPA]) ? hi : (long long)s_upre[(b + 1) * UPA])
                                     : ((hi < anchor_base + p.K1) ? hi : (anchor_base + p.K1));
    const int pos_off = (compact || banded) ? 0 : ((anchor_base == lo) ? 0 : -1);  // queue tag of the positive (k == 0) entry

    float v1c[NV], v2c[NV];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int e0 = (j + LPR * i) * VEC;
#pragma unroll
      for (int t = 0; t < VEC; t += 4) {
        const float4 a = *reinterpret_cast<const float4*>(p.v1 + (long long)b * p.D + e0 + t);
        const float4 c4 = *reinterpret_cast<const float4*>(p.v2 + (long long)b * p.D + e0 + t);
        v1c[i * VEC + t + 0] = a.x; v1c[i * VEC + t + 1] = a.y; v1c[i * VEC + t + 2] = a.z; v1c[i * VEC + t + 3] = a.w;
        v2c[i * VEC + t + 0] = c4.x; v2c[i * VEC + t + 1] = c4.y; v2c[i * VEC + t + 2] = c4.z; v2c[i * VEC + t + 3] = c4.w;
      }
    }
    // the per-row arithmetic runs on packed pairs (FFMA2): the kernel issues ~1.2 instructions per cycle and SM at the
    // headline shape, i.e. it is as close to its issue limit as to the HBM roofline, so halving the FMA count pays
    unsigned long long v1p[NV / 2], v2p[NV / 2], g1p[NV / 2], g2p[NV / 2];
#pragma unroll
    for (int n = 0; n < NV / 2; ++n) {
      v1p[n] = pack2(v1c[2 * n], v1c[2 * n + 1]);
      v2p[n] = pack2(v2c[2 * n], v2c[2 * n + 1]);
      g1p[n] = 0ull; g2p[n] = 0ull;
    }
    float ls = 0.f, lt = 0.f, se1 = 0.f, se2 = 0.f, cnt = 0.f;

    int qhead = 0, qtail = 0;

    // consume up to R*U queue entries; `avail` >= R*U unless draining
    auto consume = [&](int avail) {
      int2 ent[U];
      bool ev[U];
      uint4 w1r[U][CH], w2r[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int qi = u * R + g;
        ev[u] = qi < avail;
        ent[u] = q[(qhead + qi) & (kQueueCap - 1)];
        const long long roff = (long long)ent[u].x * p.row_stride_bytes + (long long)j * 16;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          if (ev[u]) {
            w1r[u][i] = ld16_stream(p.bank1 + roff + (long long)(LPR * i) * 16);
            w2r[u][i] = ld16_stream(p.bank2 + roff + (long long)(LPR * i) * 16);
          } else {
            w1r[u][i] = make_uint4(0u, 0u, 0u, 0u);
            w2r[u][i] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      if constexpr (kFast && sizeof(T) == 4 && kD == 128) {
        // the NEXT consume step's rows (queued already) are pulled into the L2 while this step's loads are in flight: the
        // kernel is bound by load latency at 16 warps per SM, and a prefetch holds no registers.  8 rows x (512 B of bank 1 +
        // 512 B of bank 2) = 64 lines of 128 B, two per lane
        if (p.prefetch != 0) {
          const int ln = lane & 7;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int eidx = R * U + (lane >> 3) + 4 * h;
            if (qhead + eidx < qtail) {
              const int prow = q[(qhead + eidx) & (kQueueCap - 1)].x;
              const char* pa = (ln < 4 ? p.bank1 : p.bank2) + (long long)prow * p.row_stride_bytes + (long long)(ln & 3) * 128;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
            }
          }
        }
      }
      if constexpr (kFast) {
        // ---- 16 (or 32) lanes per row, U = 4: the 8 partial dot products of a lane (4 rows x 2 directions) are reduced over
        // the row's lanes by a TRANSPOSING butterfly (8 shuffles instead of 32: every step halves the number of values a lane
        // carries), which leaves value v = (lane >> kDupShift) & 7 = (row step u, direction d) on a group of 2 (4) neighbouring
        // lanes of the row group.  exp / rcp / log then run ONCE per consume step on that one value per lane (they ran four times
        // on two values each, identically on all lanes of the row), and the eight gradient coefficients come back by 8 indexed
        // shuffles.
        unsigned long long w1p[U][NV / 2], w2p[U][NV / 2];
        float x[2 * U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float w1f[NV], w2f[NV];
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            Unpack<T>::run(w1r[u][i], &w1f[i * VEC]);
            Unpack<T>::run(w2r[u][i], &w2f[i * VEC]);
          }
          unsigned long long A1 = 0ull, A2 = 0ull;
#pragma unroll
          for (int n = 0; n < NV / 2; ++n) {
            w1p[u][n] = pack2(w1f[2 * n], w1f[2 * n + 1]);
            w2p[u][n] = pack2(w2f[2 * n], w2f[2 * n + 1]);
            ffma2(A1, w2p[u][n], v1p[n]);  // out_v1 direction: bank2 row . v1
            ffma2(A2, w1p[u][n], v2p[n]);  // out_v2 direction: bank1 row . v2
          }
          x[2 * u] = lo2(A1) + hi2(A1);
          x[2 * u + 1] = lo2(A2) + hi2(A2);
        }
        constexpr int o1 = LPR / 2, o2 = LPR / 4, o3 = LPR / 8;   // the three halving steps; the remaining ones add duplicates
        const bool hb = (lane & o1) != 0, mb = (lane & o2) != 0, lb = (lane & o3) != 0;
        float y4[4], z2[2];
#pragma unroll
        for (int k = 0; k < 4; ++k) y4[k] = (hb ? x[4 + k] : x[k]) + __shfl_xor_sync(kFull, hb ? x[k] : x[4 + k], o1);
#pragma unroll
        for (int k = 0; k < 2; ++k) z2[k] = (mb ? y4[2 + k] : y4[k]) + __shfl_xor_sync(kFull, mb ? y4[k] : y4[2 + k], o2);
        float tot = (lb ? z2[1] : z2[0]) + __shfl_xor_sync(kFull, lb ? z2[0] : z2[1], o3);
#pragma unroll
        for (int off = o3 / 2; off >= 1; off >>= 1) tot += __shfl_xor_sync(kFull, tot, off);
        // this lane's value: row step mu, direction md (0: bank-2 row . v1 -> Z1, 1: bank-1 row . v2 -> Z2)
        const int vmine = (lane >> kDupShift) & 7;
        const int mu = vmine >> 1, md = vmine & 1, mq = mu * R + g;
        const bool mvalid = mq < avail;
        const int2 ment = q[(qhead + mq) & (kQueueCap - 1)];
        const float e = ex2_approx(tot * p.k_exp);
        const float m = mvalid ? 1.f : 0.f;
        const float o = e * (md ? p.inv_Z2 : p.inv_Z1);
        const float rc = rcp_approx(o + p.c);
        const bool is_pos = (ment.y == pos_off);
        const float coef = (is_pos ? -p.c : o) * rc * (p.inv_BT * m);
        float t;
        if (is_pos) t = logf(__fdiv_rn(o, o + p.c));
        else t = -log1p_pos(fmaf(o, p.inv_mPn, p.eps_over_mPn));
        if ((lane & ((1 << kDupShift) - 1)) == 0) {   // one lane of the group that shares the value accounts for it
          const float tm = t * m, em = e * m;
          if (md == 0) { ls += tm; se1 += em; cnt += m; } else { lt += tm; se2 += em; }
          if (store_out && mvalid) (md ? p.out_v2 : p.out_v1)[lo + ment.y] = o;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float d1 = __shfl_sync(kFull, coef, ((2 * u) << kDupShift) | (lane & ~(LPR - 1) & 31));
          const float d2 = __shfl_sync(kFull, coef, ((2 * u + 1) << kDupShift) | (lane & ~(LPR - 1) & 31));
          const unsigned long long d1p = pack2(d1, d1), d2p = pack2(d2, d2);
#pragma unroll
          for (int n = 0; n < NV / 2; ++n) {
            ffma2(g1p[n], d1p, w2p[u][n]);
            ffma2(g2p[n], d2p, w1p[u][n]);
          }
        }
      } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float w1f[NV], w2f[NV];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          Unpack<T>::run(w1r[u][i], &w1f[i * VEC]);
          Unpack<T>::run(w2r[u][i], &w2f[i * VEC]);
        }
        unsigned long long w1p[NV / 2], w2p[NV / 2], A1 = 0ull, A2 = 0ull;
#pragma unroll
        for (int n = 0; n < NV / 2; ++n) {
          w1p[n] = pack2(w1f[2 * n], w1f[2 * n + 1]);
          w2p[n] = pack2(w2f[2 * n], w2f[2 * n + 1]);
          ffma2(A1, w2p[n], v1p[n]);  // out_v1 direction: bank2 row . v1
          ffma2(A2, w1p[n], v2p[n]);  // out_v2 direction: bank1 row . v2
        }
        float a1 = lo2(A1) + hi2(A1), a2 = lo2(A2) + hi2(A2);
#pragma unroll
        for (int off = LPR / 2; off >= 1; off >>= 1) {
          a1 += __shfl_xor_sync(kFull, a1, off);
          a2 += __shfl_xor_sync(kFull, a2, off);
        }
        const float e1 = ex2_approx(a1 * p.k_exp);
        const float e2 = ex2_approx(a2 * p.k_exp);
        const float m = ev[u] ? 1.f : 0.f;
        se1 = fmaf(e1, m, se1);
        se2 = fmaf(e2, m, se2);
        cnt += m;
        if constexpr (FULL) {
          const bool is_pos = (ent[u].y == pos_off);
          const float o1 = e1 * p.inv_Z1, o2 = e2 * p.inv_Z2;
          const float rc1 = rcp_approx(o1 + p.c), rc2 = rcp_approx(o2 + p.c);
          const float sc = p.inv_BT * m;
          const float d1 = (is_pos ? -p.c : o1) * rc1 * sc;
          const float d2 = (is_pos ? -p.c : o2) * rc2 * sc;
          float t1, t2;
          if (is_pos) {
            t1 = logf(__fdiv_rn(o1, o1 + p.c));
            t2 = logf(__fdiv_rn(o2, o2 + p.c));
          } else {
            t1 = -log1p_pos(fmaf(o1, p.inv_mPn, p.eps_over_mPn));
            t2 = -log1p_pos(fmaf(o2, p.inv_mPn, p.eps_over_mPn));
          }
          ls = fmaf(t1, m, ls);
          lt = fmaf(t2, m, lt);
          const unsigned long long d1p = pack2(d1, d1), d2p = pack2(d2, d2);
#pragma unroll
          for (int n = 0; n < NV / 2; ++n) {
            ffma2(g1p[n], d1p, w2p[n]);
            ffma2(g2p[n], d2p, w1p[n]);
          }
          if (store_out && ev[u] && j == 0) {
            p.out_v1[lo + ent[u].y] = o1;
            p.out_v2[lo + ent[u].y] = o2;
          }
        } else {
          if (store_out && ev[u] && j == 0) {
            p.out_v1[lo + ent[u].y] = e1;
            p.out_v2[lo + ent[u].y] = e2;
          }
        }
      }
          }
    };

    auto fetch_idx = [&](long long pos) -> long long { return contrast_entry(p, pos, b, anchor_base); };
    if (banded) {
      // warp i of the anchor's wpu warps takes blocks g = i, i + wpu, ... of the sequence g = blk * NC + c (32-entry block blk of
      // unit c): every unit list is band-sorted, so the sequence sweeps the bank once, and the anchor's entries are spread
      // over its warps to within one block whatever NC and the unit lengths are
      const int i = (int)(gw - (long long)b * p.wpu);
      const int ncnt = lane < p.NC ? p.ucount[b * p.NC + lane] : 0;   // (NC <= 32: one unit length per lane)
      int nmax = ncnt;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) nmax = max(nmax, __shfl_xor_sync(kFull, nmax, off));
      const int G = ((nmax + 31) >> 5) * p.NC;
      const int* src = p.cl + (size_t)b * p.NC * kFilterChunk;
      if (i == 0) {   // the anchor's positive (k = 0) rides with its first warp
        const long long r = contrast_entry(p, anchor_base, b, anchor_base);
        if (r >= p.row_begin && r < p.row_end) {
          if (lane == 0) q[qtail & (kQueueCap - 1)] = make_int2((int)(r - p.row_begin), 0);
          qtail += 1;
        }
      }
      auto fetch_block = [&](int gg, int& cnt_blk) -> int {
        if (gg >= G) { cnt_blk = 0; return 0; }
        const int blk = gg / p.NC, c = gg - blk * p.NC;
        const int n = __shfl_sync(kFull, ncnt, c), off = blk * 32;
        cnt_blk = n - off < 0 ? 0 : (n - off < 32 ? n - off : 32);
        return lane < cnt_blk ? src[(size_t)c * kFilterChunk + off + lane] : 0;
      };
      int g2 = i, cnt_nxt;
      int nxt = fetch_block(g2, cnt_nxt);   // the next block's indices are loaded before this block's rows are consumed
#pragma unroll 1
      while (g2 < G) {
        const int cur_row = nxt, cnt_blk = cnt_nxt;
        g2 += p.wpu;
        nxt = fetch_block(g2, cnt_nxt);
        if (lane < cnt_blk) q[(qtail + lane) & (kQueueCap - 1)] = make_int2(cur_row, 1);
        qtail += cnt_blk;
        __syncwarp();
        while (qtail - qhead >= R * U) {
          consume(R * U);
          qhead += R * U;
        }
        __syncwarp();
      }
    } else if (compact) {
      // entries of this segment, unit by unit: every one of them is scored (the filter kernel dropped the rest)
      long long cur = lo;
      while (cur < seg_hi) {
        while ((long long)s_upre[cu + 1] <= cur) ++cu;
        const int uc = cu - b * UPA;
        const long long uend = (seg_hi < (long long)s_upre[cu + 1]) ? seg_hi : (long long)s_upre[cu + 1];
        const int off = (int)(cur - s_upre[cu]), n = (int)(uend - cur);
        const int* src = p.cl + ((size_t)b * p.NC + (uc > 0 ? uc - 1 : 0)) * kFilterChunk + off;
#pragma unroll 1
        for (int i = 0; i < n; i += 32) {
          if (i + lane < n) {
            int row;
            if (uc == 0) row = (int)(contrast_entry(p, anchor_base, b, anchor_base) - p.row_begin);
            else row = src[i + lane];
            q[(qtail + lane) & (kQueueCap - 1)] = make_int2(row, uc == 0 ? 0 : 1);
          }
          qtail += (n - i < 32) ? (n - i) : 32;
          __syncwarp();
          while (qtail - qhead >= R * U) {
            consume(R * U);
            qhead += R * U;
          }
          __syncwarp();
        }
        cur = uend;
      }
    } else {
    long long base = lo;
    // a scan step covers 64 entries (two per lane: half the loop iterations, ballots and queue bookkeeping per entry -- the
    // scan is pure instruction overhead on a shard that owns 1/R of the rows); the index loads run two steps ahead of
    // their use in a rotating register window (not an unrolled loop: the body contains the whole consume step, and four
    // copies of it made the kernel instruction-fetch bound, 0.43 -> 0.63 ms at the headline shape)
    auto fetch_local = [&](long long pos) -> int {   // local row of the entry, -1 when another shard owns it (rows < 2^31)
      if (pos >= seg_hi) return -1;
      const long long r = fetch_idx(pos);
      return (r >= p.row_begin && r < p.row_end) ? (int)(r - p.row_begin) : -1;
    };
    int ra0 = fetch_local(base + 2 * lane), rb0 = fetch_local(base + 2 * lane + 1);
    int ra1 = fetch_local(base + 64 + 2 * lane), rb1 = fetch_local(base + 64 + 2 * lane + 1);
#pragma unroll 1
    while (base < seg_hi) {
      const int ra = ra0, rb = rb0;
      const long long pidx = base + 2 * lane;
      ra0 = ra1; rb0 = rb1;
      ra1 = fetch_local(pidx + 128);
      rb1 = fetch_local(pidx + 129);
      const unsigned ma = __ballot_sync(kFull, ra >= 0), mb = __ballot_sync(kFull, rb >= 0);
      const unsigned below = (1u << lane) - 1u;
      if (ra >= 0) q[(qtail + __popc(ma & below)) & (kQueueCap - 1)] = make_int2(ra, (int)(pidx - lo));
      else if (store_out && pidx < seg_hi) { p.out_v1[pidx] = 0.f; p.out_v2[pidx] = 0.f; }
      if (rb >= 0) q[(qtail + __popc(ma) + __popc(mb & below)) & (kQueueCap - 1)] = make_int2(rb, (int)(pidx + 1 - lo));
      else if (store_out && pidx + 1 < seg_hi) { p.out_v1[pidx + 1] = 0.f; p.out_v2[pidx + 1] = 0.f; }
      qtail += __popc(ma) + __popc(mb);
      __syncwarp();
      while (qtail - qhead >= R * U) {
        consume(R * U);
        qhead += R * U;
      }
      __syncwarp();
      base += 64;
    }
    }
    while (qtail - qhead > 0) {
      const int avail = qtail - qhead;
      consume(avail);
      qhead += (avail < R * U) ? avail : R * U;
    }
    __syncwarp();

    // ---- flush this (warp, anchor) partial ----
    float g1[NV], g2[NV];
#pragma unroll
    for (int n = 0; n < NV / 2; ++n) {
      g1[2 * n] = lo2(g1p[n]); g1[2 * n + 1] = hi2(g1p[n]);
      g2[2 * n] = lo2(g2p[n]); g2[2 * n + 1] = hi2(g2p[n]);
    }
#pragma unroll
    for (int off = 16; off >= LPR; off >>= 1) {
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        g1[n] += __shfl_xor_sync(kFull, g1[n], off);
        g2[n] += __shfl_xor_sync(kFull, g2[n], off);
      }
      ls += __shfl_xor_sync(kFull, ls, off);
      lt += __shfl_xor_sync(kFull, lt, off);
      se1 += __shfl_xor_sync(kFull, se1, off);
      se2 += __shfl_xor_sync(kFull, se2, off);
      cnt += __shfl_xor_sync(kFull, cnt, off);
    }
    if constexpr (kFast) {   // the scalar sums are spread over the lanes of a row group (one value per lane pair)
#pragma unroll
      for (int off = LPR / 2; off >= 1; off >>= 1) {
        ls += __shfl_xor_sync(kFull, ls, off);
        lt += __shfl_xor_sync(kFull, lt, off);
        se1 += __shfl_xor_sync(kFull, se1, off);
        se2 += __shfl_xor_sync(kFull, se2, off);
        cnt += __shfl_xor_sync(kFull, cnt, off);
      }
    }
    float* slot = cta_red ? red_smem + (warp * 2 + (b - b_first_cta)) * kSW
                : banded ? p.slots + (long long)gw * (2 * p.D + kSlotExtra)
                : compact ? p.slots + ((long long)gw + b) * (2 * p.D + kSlotExtra)
                          : p.slots + ((long long)gw * p.maxseg + seg) * (2 * p.D + kSlotExtra);
    if (g == 0) {
      if constexpr (FULL) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int e0 = (j + LPR * i) * VEC;
#pragma unroll
          for (int t = 0; t < VEC; t += 4) {
            *reinterpret_cast<float4*>(slot + e0 + t) =
                make_float4(g1[i * VEC + t], g1[i * VEC + t + 1], g1[i * VEC + t + 2], g1[i * VEC + t + 3]);
            *reinterpret_cast<float4*>(slot + p.D + e0 + t) =
                make_float4(g2[i * VEC + t], g2[i * VEC + t + 1], g2[i * VEC + t + 2], g2[i * VEC + t + 3]);
          }
        }
      }
      if (j == 0) {
        *reinterpret_cast<float4*>(slot + 2 * p.D) = make_float4(ls, lt, se1, se2);
        *reinterpret_cast<float4*>(slot + 2 * p.D + 4) = make_float4(cnt, 0.f, 0.f, 0.f);
      }
    }
    lo = seg_hi;
    ++seg;
  }
  if (cta_red) {
    // fold the 8 warps' partials in fixed order (bit-reproducible) and write this CTA's two slots
    __syncthreads();
    float* out = p.slots + (long long)blockIdx.x * 2 * kSW;
    for (int i = threadIdx.x; i < 2 * kSW; i += kThreads) {
      const int sg = i / kSW, col = i - sg * kSW;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += red_smem[(w * 2 + sg) * kSW + col];
      out[i] = acc;
    }
  }
}

// One CTA per anchor: fixed-order sum of the warp partials that intersect the anchor's K1 pairs; the last
// CTA to finish folds the per-anchor scalars into result[] (also in fixed order) -> bit-reproducible.
// The slot offsets are computed once per CTA into shared memory (no 64-bit divisions in the summation loop).
constexpr int kFinalizeThreads = 1024;
constexpr int kFinalizeList = 1024;
constexpr int kFinalizeScratch = 4096;  // doubles

__device__ __forceinline__ void finalize_body(const FinalizeParams& f, const int b) {
  __shared__ long long s_off[kFinalizeList];
  __shared__ double s_part[kFinalizeScratch];
  __shared__ bool is_last;
  __shared__ uint32_t s_epoch;
  if (f.xchg && threadIdx.x == 0)   // advanced by the last block only, i.e. after every block has read it
    s_epoch = *reinterpret_cast<volatile uint32_t*>(f.peers.buf[f.rank] + f.off_ctl + 8) + 1u;
  // pair space: B*K1 entries, anchor b = [b K1, (b+1) K1); compact mode: the entries this shard owns, anchor b =
  // [anchor_start[b], anchor_start[b+1]) as the scoring pass measured them
  const bool compact = f.anchor_start != nullptr;
  const long long P = compact ? f.anchor_start[f.B] : (long long)f.B * f.K1;
  const long long p0 = compact ? f.anchor_start[b] : (long long)b * f.K1;
  const long long p1 = compact ? f.anchor_start[b + 1] : p0 + f.K1;
  // slot-writing units: single warps (group = 1) or whole CTAs of `group` warps; unit u covers pairs [bnd(u), bnd(u+1))
  const long long G = f.group, NU = (f.NW + G - 1) / G;
  auto bnd = [&](long long u) -> long long {
    const long long w = u * G < f.NW ? u * G : f.NW;
    return (P * w) / f.NW;
  };
  // candidate units: the ones around p0 * NU / P .. p1 * NU / P, one extra on either side (the integer floors of bnd() can be
  // off by one); units that do not overlap the anchor's range are dropped by the `valid` test below.  Every thread
  // computes the same bounds: no serial search by one thread, no barrier
  long long first = 1, last = 0;   // (compact mode) p1 <= p0: nothing of this anchor lives in this shard
  if (f.wpa > 0) {
    first = (long long)b * f.wpa;
    last = first + f.wpa - 1;
  } else if (p1 > p0) {
    first = ((p0 * f.NW) / P) / G - 1;
    if (first < 0) first = 0;
    last = ((p1 * f.NW + P - 1) / P) / G + 1;
    if (last >= NU) last = NU - 1;
    if (first > last) first = last;
  }
  const int slot_w = 2 * f.D + kSlotExtra;
  const int ncol4 = slot_w / 4;                       // float4 columns per slot (grad_v1 | grad_v2 | scalars)
  int parts = kFinalizeThreads / ncol4;               // entry-parallel groups of ncol4 threads
  if (parts > kFinalizeScratch / slot_w) parts = kFinalizeScratch / slot_w;
  const int part = threadIdx.x / ncol4, c4 = threadIdx.x % ncol4;
  const bool active = part < parts;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (long long chunk = first; chunk <= last; chunk += kFinalizeList) {
    const int n = (int)((last - chunk + 1 < kFinalizeList) ? (last - chunk + 1) : kFinalizeList);
    for (int i = threadIdx.x; i < n; i += kFinalizeThreads) {
      const long long w = chunk + i;
      const long long lo = bnd(w);
      const long long hi = bnd(w + 1);
      const bool valid = f.wpa > 0 || (hi > p0 && hi > lo && lo < p1);
      s_off[i] = !valid ? -1 : f.wpa > 0 ? w * (long long)slot_w : compact ? (w + b) * (long long)slot_w
                                       : (w * f.maxseg + (b - (int)(lo / f.K1))) * (long long)slot_w;
    }
    __syncthreads();
    if (active) {
      // entries part, part+parts, ...: independent 128-bit loads, all in flight together
#pragma unroll 8
      for (int i = part; i < n; i += parts) {
        const long long off = s_off[i];
        if (off >= 0) {
          const float4 v = *reinterpret_cast<const float4*>(f.slots + off + 4 * c4);
          a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
        }
      }
    }
    __syncthreads();
  }
  if (active) {
    double* dst = s_part + part * slot_w + 4 * c4;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  }
  __syncthreads();
  const int ncols = 2 * f.D;
  for (int col = threadIdx.x; col < ncols + 5; col += kFinalizeThreads) {
    double acc = 0.0;
    for (int q = 0; q < parts; ++q) acc += s_part[q * slot_w + col];  // fixed order -> deterministic
    if (f.xchg) {
      const uint32_t e = s_epoch;
      const size_t par = (size_t)(e & 1u) * f.parity_stride + f.off_slots;
      const size_t BD = (size_t)f.B * f.D;
      const size_t word = col < f.D ? (size_t)b * f.D + col
                        : col < ncols ? BD + (size_t)b * f.D + (col - f.D)
                                      : 2 * BD + 8 + 8 * (size_t)b + (col - ncols);
      const uint32_t bits = __float_as_uint((float)acc);
      for (int r = 0; r < f.world; ++r)
        p2p::ll_store(f.peers.buf[r] + par + ((size_t)f.rank * f.slot_words + word) * 8, bits, e);
      const char* mine = f.peers.buf[f.rank] + par;
      float sum = __uint_as_float(p2p::ll_load(mine + word * 8, e, f.timeout));
      for (int r = 1; r < f.world; ++r)   // rank order: the same bits on every rank
        sum += __uint_as_float(p2p::ll_load(mine + ((size_t)r * f.slot_words + word) * 8, e, f.timeout));
      if (col < ncols) f.reduced[word] = sum;
      else f.anchor_part[b * 8 + (col - ncols)] = (double)sum;
      continue;
    }
    if (col < ncols) {
      if (f.full) {
        if (col < f.D) f.grad_v1[(long long)b * f.D + col] = (float)acc;
        else f.grad_v2[(long long)b * f.D + (col - f.D)] = (float)acc;
      }
    } else {
      f.anchor_part[b * 8 + (col - ncols)] = acc;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(f.ticket, 1u);
    is_last = (t == (unsigned)(f.B - 1));
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // the per-anchor scalars come into shared memory with ONE load per thread (a chain of B dependent-latency loads by five
    // threads cost ~3 us); the sums below keep their fixed order
    const bool staged = f.B * 8 <= kFinalizeScratch;
    if (staged) {
      __syncthreads();   // s_part is free: every thread has left the column sums above
      for (int i = threadIdx.x; i < f.B * 8; i += kFinalizeThreads) s_part[i] = __ldcg(&f.anchor_part[i]);
      __syncthreads();
    }
    auto ap = [&](int i) { return staged ? s_part[i] : __ldcg(&f.anchor_part[i]); };
    if (threadIdx.x < 5) {
      double s = 0.0;
      for (int a = 0; a < f.B; ++a) s += ap(a * 8 + threadIdx.x);
      f.result[threadIdx.x] = (threadIdx.x < 2) ? (f.full ? -s / (double)f.B : 0.0) : s;
    } else if (threadIdx.x == 5) {  // total loss, also as float32 (slot 6) so the caller needs no cast kernel
      double s0 = 0.0, s1 = 0.0;
      for (int a = 0; a < f.B; ++a) { s0 += ap(a * 8); s1 += ap(a * 8 + 1); }
      const double tot = f.full ? (-s0 / (double)f.B) + (-s1 / (double)f.B) : 0.0;
      f.result[5] = tot;
      f.result[6] = 0.0;
      reinterpret_cast<float*>(&f.result[6])[0] = (float)tot;
      if (f.xchg) f.reduced[2 * (size_t)f.B * f.D + 5] = (float)tot;
    } else if (threadIdx.x == 7) {
      f.result[7] = 0.0;
      if (f.xchg)   // every block of this rank has consumed epoch e: publish it (the next exchange uses e + 1)
        *reinterpret_cast<volatile uint32_t*>(f.peers.buf[f.rank] + f.off_ctl + 8) = s_epoch;
    }
    if (f.xchg && threadIdx.x < 5) f.reduced[2 * (size_t)f.B * f.D + threadIdx.x] = (float)f.result[threadIdx.x];
  }
}

struct UpdateParams {
  char* bank1;
  char* bank2;
  long long row_stride_bytes;
  const float* v1;
  const float* v2;
  const long long* y;
  int B, D;
  long long row_begin, row_end;
  float m, om;
};

// Momentum update: one warp per (anchor, bank).  Canonical order: element e -> lane (e/4)%32, per-lane
// fmaf fold in increasing e, xor butterfly 16..1 (bit-identical to oracle/crd_oracle.c canonical_sumsq).
template <typename T>
__device__ __forceinline__ void update_body(const UpdateParams& u, const int wid) {
  const int lane = threadIdx.x & 31;
  if (wid >= 2 * u.B) return;
  const int b = wid >> 1;
  const int which = wid & 1;
  const long long* y = u.y;
  const int B = u.B, D = u.D;
  const long long row_begin = u.row_begin, row_end = u.row_end, row_stride_bytes = u.row_stride_bytes;
  char* bank1 = u.bank1;
  char* bank2 = u.bank2;
  const float* v1 = u.v1;
  const float* v2 = u.v2;
  const float m = u.m, om = u.om;
  const long long r = y[b];
  if (r < row_begin || r >= row_end) return;
  bool dup = false;
  for (int b2 = b + 1 + lane; b2 < B; b2 += 32) dup |= (y[b2] == r);
  if (__any_sync(0xffffffffu, dup)) return;  // a later occurrence of the same index wins
  char* rowp = (which ? bank2 : bank1) + (r - row_begin) * row_stride_bytes;
  const float* v = (which ? v2 : v1) + (long long)b * D;
  constexpr int kMaxQ = 8;  // D <= 1024
  float4 pv[kMaxQ];
  float acc = 0.f;
  const int nq = D >> 2;
#pragma unroll
  for (int t = 0; t < kMaxQ; ++t) {
    const int qd = lane + 32 * t;
    if (qd < nq) {
      float4 mem;
      if constexpr (sizeof(T) == 4) {
        mem = *reinterpret_cast<const float4*>(rowp + (long long)qd * 16);
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(rowp + (long long)qd * 8);
        mem = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                          __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
      }
      const float4 vv = *reinterpret_cast<const float4*>(v + qd * 4);
      float4 pp;
      pp.x = __fadd_rn(__fmul_rn(m, mem.x), __fmul_rn(om, vv.x));
      pp.y = __fadd_rn(__fmul_rn(m, mem.y), __fmul_rn(om, vv.y));
      pp.z = __fadd_rn(__fmul_rn(m, mem.z), __fmul_rn(om, vv.z));
      pp.w = __fadd_rn(__fmul_rn(m, mem.w), __fmul_rn(om, vv.w));
      acc = __fmaf_rn(pp.x, pp.x, acc);
      acc = __fmaf_rn(pp.y, pp.y, acc);
      acc = __fmaf_rn(pp.z, pp.z, acc);
      acc = __fmaf_rn(pp.w, pp.w, acc);
      pv[t] = pp;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
  const float nrm = __fsqrt_rn(acc);
#pragma unroll
  for (int t = 0; t < kMaxQ; ++t) {
    const int qd = lane + 32 * t;
    if (qd < nq) {
      float4 o;
      o.x = __fdiv_rn(pv[t].x, nrm); o.y = __fdiv_rn(pv[t].y, nrm);
      o.z = __fdiv_rn(pv[t].z, nrm); o.w = __fdiv_rn(pv[t].w, nrm);
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(rowp + (long long)qd * 16) = o;
      } else {
        const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o.x, o.y);
        const __nv_bfloat162 hi2 = __floats2bfloat162_rn(o.z, o.w);
        uint2 u;
        u.x = *reinterpret_cast<const unsigned*>(&lo2);
        u.y = *reinterpret_cast<const unsigned*>(&hi2);
        *reinterpret_cast<uint2*>(rowp + (long long)qd * 8) = u;
      }
    }
  }
}

__global__ void __launch_bounds__(kFinalizeThreads) crd_finalize_kernel(const FinalizeParams f) {
  finalize_body(f, blockIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(kFinalizeThreads) crd_update_kernel(const UpdateParams u) {
  update_body<T>(u, blockIdx.x * (kFinalizeThreads / 32) + (threadIdx.x >> 5));
}

// finalize (blocks [0,B)) and momentum update (remaining blocks) in one launch: both only depend on the score
// kernel having finished, and touch disjoint memory.
template <typename T>
__global__ void __launch_bounds__(kFinalizeThreads) crd_finalize_update_kernel(const FinalizeParams f, const UpdateParams u) {
  if ((int)blockIdx.x < f.B) finalize_body(f, blockIdx.x);
  else update_body<T>(u, ((int)blockIdx.x - f.B) * (kFinalizeThreads / 32) + (threadIdx.x >> 5));
}

#include "crd_stream.cuh"
#include "crd_tc_stream.cuh"

// ------------------------------------------------------------------------------------------------------
// Alias-method draw: one Philox4x32-10 block per output (identical stream to oracle/crd_oracle.c).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alias_draw_kernel(const float* __restrict__ prob,
                                                         const long long* __restrict__ alias, long long n,
                                                         long long count, unsigned long long seed,
                                                         unsigned long long offset, const long long* __restrict__ y,
                                                         long long K1, long long row_base, long long* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool small = count < (1ll << 32) && K1 < (1ll << 32);  // 32-bit division: the 64-bit one is a ~100-instruction call
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    if (y != nullptr) {
      long long row;
      bool first;
      if (small) {
        const unsigned q = (unsigned)i / (unsigned)K1;
        row = q;
        first = (unsigned)i - q * (unsigned)K1 == 0u;
      } else {
        row = i / K1;
        first = i - row * K1 == 0;
      }
      if (first) {
        out[i] = y[row];
        continue;
      }
    }
    unsigned r[4];
    philox4x32_10(seed, offset + (unsigned long long)i, r);
    const unsigned long long bits = ((unsigned long long)r[0] << 32) | (unsigned long long)r[1];
    const long long kk = (long long)__umul64hi(bits, (unsigned long long)n);
    const float u = (float)(r[2] >> 8) * 5.9604644775390625e-08f;
    out[i] = row_base + ((prob == nullptr || u < prob[kk]) ? kk : alias[kk]);  // prob == NULL: uniform tables (prob = 1 everywhere)
  }
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
using ScoreKernel = void (*)(const ScoreParams);
struct Variant {
  ScoreKernel full, sums;
  int bps;
};

template <typename T, int LPR, int CH, int U, int BPS>
static Variant make_variant() {
  return Variant{crd_score_kernel<T, LPR, CH, U, BPS, true>, crd_score_kernel<T, LPR, CH, U, BPS, false>, BPS};
}

// (dtype, D, variant) -> kernel.  variant 0 is the tuned default for that D.
static bool pick_variant(int dtype, int D, int variant, Variant* out) {
  if (dtype == CRDPN_F32) {
    switch (D) {
      case 32:
        *out = make_variant<float, 8, 1, 4, 2>(); return variant == 0;
      case 64:
        if (variant == 0 || variant == 1) { *out = make_variant<float, 8, 2, 4, 2>(); return true; }
        if (variant == 2) { *out = make_variant<float, 16, 1, 4, 2>(); return true; }
        return false;
      case 128:
        if (variant == 0 || variant == 1) { *out = make_variant<float, 16, 2, 4, 2>(); return true; }
        if (variant == 2) { *out = make_variant<float, 8, 4, 2, 1>(); return true; }
        if (variant == 3) { *out = make_variant<float, 32, 1, 4, 2>(); return true; }
        if (variant == 4) { *out = make_variant<float, 32, 1, 8, 2>(); return true; }
        if (variant == 5) { *out = make_variant<float, 16, 2, 2, 3>(); return true; }
        if (variant == 6) { *out = make_variant<float, 8, 4, 2, 2>(); return true; }
        if (variant == 7) { *out = make_variant<float, 16, 2, 8, 1>(); return true; }
        return false;
      case 256:
        if (variant == 0 || variant == 1) { *out = make_variant<float, 32, 2, 4, 2>(); return true; }
        if (variant == 2) { *out = make_variant<float, 16, 4, 2, 1>(); return true; }
        return false;
      case 512:
        *out = make_variant<float, 32, 4, 2, 1>(); return variant == 0;
      default: return false;
    }
  } else if (dtype == CRDPN_BF16) {
    switch (D) {
      case 64:
        *out = make_variant<__nv_bfloat16, 8, 1, 4, 2>(); return variant == 0;
      case 128:
        // round 1: 256-byte rows are transaction-bound, 8 row-steps in flight (0.367 ms at the headline config) beat the fp32
        // kernel's shape (0.506 ms).  Round 2: with the transposed score reduction (U = 4 only) the kernel is 39 % shorter in
        // instructions and the 4-step shape wins: 0.418 -> 0.287 ms, B = 138: 1.21 -> 0.85 ms (profiles/r2_bf16_gather_ab.py)
        if (variant == 0 || variant == 4) { *out = make_variant<__nv_bfloat16, 16, 1, 4, 2>(); return true; }
        if (variant == 2) { *out = make_variant<__nv_bfloat16, 16, 1, 8, 2>(); return true; }
        if (variant == 1) { *out = make_variant<__nv_bfloat16, 8, 2, 4, 2>(); return true; }
        if (variant == 3) { *out = make_variant<__nv_bfloat16, 4, 4, 2, 2>(); return true; }
        return false;
      case 256:
        if (variant == 0 || variant == 1) { *out = make_variant<__nv_bfloat16, 16, 2, 4, 2>(); return true; }
        if (variant == 2) { *out = make_variant<__nv_bfloat16, 8, 4, 2, 1>(); return true; }
        return false;
      case 512:
        *out = make_variant<__nv_bfloat16, 32, 2, 4, 2>(); return variant == 0;
      default: return false;
    }
  }
  return false;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int maxseg_for(long long P, long long NW, long long K1) {
  const long long len_max = (P + NW - 1) / NW;
  return (int)((len_max + K1 - 1) / K1 + 1);
}

static inline long long filter_chunks(long long K1) { return (K1 - 1 + kFilterChunk - 1) / kFilterChunk; }
static size_t workspace_slots_end(long long B, long long K1, long long D, long long NW) {
  const long long P = B * K1;
  const size_t head = 16 + align_up((size_t)B * 8 * sizeof(double), 16);
  const size_t slots = (size_t)NW * (size_t)maxseg_for(P, NW, K1) * (size_t)(2 * D + kSlotExtra) * sizeof(float);
  return align_up(head + slots, 256);
}
// behind the slots: the compact lists of the row-sharded step (anchor_start [B+1] i64 | ucount [B*NC] i32 | cl [B*NC*chunk] i32)
static size_t workspace_bytes_for(long long B, long long K1, long long D, long long NW) {
  const long long NC = filter_chunks(K1);
  return workspace_slots_end(B, K1, D, NW) + align_up((size_t)(B + 1) * 8, 256) + align_up((size_t)(B * NC) * 4, 256) +
         (size_t)(B * NC) * kFilterChunk * 4;
}

}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_crd_workspace_bytes(int64_t B, int64_t K1, int64_t D, int device, size_t* bytes) {
  if (!bytes || B <= 0 || K1 <= 0 || D <= 0) return fail(CRDPN_E_BADARG, "crdpn_crd_workspace_bytes: bad argument");
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  *bytes = workspace_bytes_for(B, K1, D, (long long)di.sms * kMaxBlocksPerSM * kWarps);
  return CRDPN_OK;
}

static int check_update_args(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                             const float* v1, const float* v2, const int64_t* y, int64_t B, int64_t D,
                             int64_t row_begin, int64_t row_end) {
  if (!v1 || !v2 || !y) return fail(CRDPN_E_BADARG, "crdpn momentum update: null pointer");
  if (B <= 0 || D <= 0 || row_end < row_begin) return fail(CRDPN_E_BADARG, "crdpn momentum update: bad size");
  if (row_end > row_begin && (!bank1 || !bank2)) return fail(CRDPN_E_BADARG, "crdpn momentum update: null bank");
  if (D % 4 != 0 || D > 1024) return fail(CRDPN_E_UNSUPPORTED, "crdpn momentum update: feat_dim must be a multiple of 4, <= 1024");
  if (bank_dtype != CRDPN_F32 && bank_dtype != CRDPN_BF16) return fail(CRDPN_E_UNSUPPORTED, "crdpn momentum update: bank dtype");
  const size_t esz = (bank_dtype == CRDPN_BF16) ? 2 : 4;
  if (((uintptr_t)bank1 | (uintptr_t)bank2 | (uintptr_t)v1 | (uintptr_t)v2) & 15 || ((size_t)row_stride * esz) % 16 != 0)
    return fail(CRDPN_E_ALIGN, "crdpn momentum update: 16-byte alignment required");
  return CRDPN_OK;
}

static UpdateParams make_update_params(void* bank1, void* bank2, int64_t row_stride, int bank_dtype, const float* v1,
                                       const float* v2, const int64_t* y, int64_t B, int64_t D, int64_t row_begin,
                                       int64_t row_end, float m, float om) {
  UpdateParams u;
  u.bank1 = (char*)bank1; u.bank2 = (char*)bank2;
  u.row_stride_bytes = (long long)row_stride * ((bank_dtype == CRDPN_BF16) ? 2 : 4);
  u.v1 = v1; u.v2 = v2; u.y = (const long long*)y;
  u.B = (int)B; u.D = (int)D; u.row_begin = row_begin; u.row_end = row_end; u.m = m; u.om = om;
  return u;
}

// how the scoring pass obtains contrast_idx when it is not an int64 list (ScoreParams::idx_mode)
struct IdxSource {
  int mode;                 // 0: int64 list (score_impl's contrast_idx); 1: int32 list; 2: drawn inside the kernel (uniform tables)
  const int* idx32;
  const int64_t* y;
  uint64_t seed, offset;
  int64_t draw_n, draw_base;
  int compact;              // 1: run crd_shard_filter_kernel first and score the compact lists (y must be set); 2: the caller has
                            // launched it; 3 / 4: the same with band-sorted lists (crd_band_sort_kernel, ScoreParams::compact == 3)
  int all_in_shard;         // hint: every entry lives in this shard (in-shard negatives): the band sort skips its survivor compaction
};

// compact lists of the row-sharded step inside the workspace (behind the slots)
struct CompactPlan {
  long long NC, NW;
  long long* anchor_start;
  int* ucount;
  int* cl;
};
// warps per anchor of the band-sorted mode, 0 = not possible (too few rows to band, more than 32 chunks per anchor: the unit
// lengths ride one per lane, or more anchors than resident warps)
static long long banded_wpu(long long B, long long NC, long long NW, long long rows_local) {
  if (rows_local < 4096 || B < 1 || NC < 1 || NC > 32 || B > NW) return 0;
  return NW / B;
}

static bool plan_compact(int bank_dtype, int64_t D, int variant, int64_t B, int64_t K1, int64_t row_begin, int64_t row_end,
                         int sms, void* workspace, CompactPlan* out) {
  Variant var;
  if ((variant & 0x100) || !pick_variant(bank_dtype, (int)D, variant & 0x1f, &var)) return false;
  const long long NW = (long long)sms * var.bps * kWarps, NC = filter_chunks(K1);
  if (NC < 1 || (B * (NC + 1) >= kMaxUnits && !(variant & 0x400)) || B > NW || row_end <= row_begin) return false;
  char* cbase = (char*)workspace + workspace_slots_end(B, K1, D, NW);
  out->NC = NC; out->NW = NW;
  out->anchor_start = (long long*)cbase;
  out->ucount = (int*)(cbase + align_up((size_t)(B + 1) * 8, 256));
  out->cl = (int*)((char*)out->ucount + align_up((size_t)(B * NC) * 4, 256));
  return true;
}
// a second stream per device so that the filter pre-pass runs beside the anchors' all-gather (fork / join by events;
// inside a stream capture the pair becomes two parallel branches of the graph)
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static int side_stream(SideStream** out) {
  static SideStream side[64];
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  SideStream& x = side[device & 63];
  if (!x.s) {
    CRDPN_CUDA(cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking));
    CRDPN_CUDA(cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming));
    CRDPN_CUDA(cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming));
  }
  *out = &x;
  return CRDPN_OK;
}

// score (+ finalize); when `upd` is given the momentum update rides in the finalize launch.
static int score_impl(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                      const float* v1, const float* v2, const int64_t* contrast_idx,
                      int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
                      float T, float Z1, float Z2, float eps, float* out_v1, float* out_v2, double* result,
                      float* grad_v1, float* grad_v2, void* workspace, size_t workspace_bytes, int variant,
                      const UpdateParams* upd, void* stream, const FinalizeParams* xchg = nullptr,
                      const IdxSource* src = nullptr) {
  if (!v1 || !v2 || (!contrast_idx && (!src || src->mode == 0)) || !result || !workspace)
    return fail(CRDPN_E_BADARG, "crdpn_crd_score: null pointer");
  if (src && ((src->mode == 1 && !src->idx32) || (src->mode == 2 && (!src->y || src->draw_n <= 0)) || src->mode < 0 || src->mode > 2 ||
              (src->mode == 0 && !contrast_idx)))
    return fail(CRDPN_E_BADARG, "crdpn_crd_score: bad index source");
  if (src && (out_v1 || out_v2) && src->mode == 2)
    return fail(CRDPN_E_BADARG, "crdpn_crd_score: per-entry outputs need a materialised contrast_idx");
  if (B <= 0 || K1 <= 0 || D <= 0 || n_data <= 0 || row_end < row_begin || !(T > 0.f))
    return fail(CRDPN_E_BADARG, "crdpn_crd_score: bad size");
  if (row_end > row_begin && (!bank1 || !bank2)) return fail(CRDPN_E_BADARG, "crdpn_crd_score: null bank");
  if ((out_v1 == nullptr) != (out_v2 == nullptr)) return fail(CRDPN_E_BADARG, "crdpn_crd_score: out_v1/out_v2 must both be set or both null");
  if (B * K1 >= (1ll << 31) || row_end - row_begin >= (1ll << 31))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_score: B*K1 and local rows must be < 2^31");
  const bool full = (Z1 > 0.f && Z2 > 0.f);
  if (full && (!grad_v1 || !grad_v2)) return fail(CRDPN_E_BADARG, "crdpn_crd_score: null grad buffer");
  const size_t esz = (bank_dtype == CRDPN_BF16) ? 2 : 4;
  if (bank_dtype != CRDPN_F32 && bank_dtype != CRDPN_BF16) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_score: bank dtype");
  if (((uintptr_t)bank1 | (uintptr_t)bank2 | (uintptr_t)v1 | (uintptr_t)v2 | (uintptr_t)workspace) & 15 ||
      ((size_t)row_stride * esz) % 16 != 0 || (D * 4) % 16 != 0)
    return fail(CRDPN_E_ALIGN, "crdpn_crd_score: banks, embeddings, workspace and row pitch must be 16-byte aligned");
  Variant var;
  // bit 8 of `variant`: anchor-aligned partition -- every anchor's list is cut into the same number S of slices, one per
  // warp (B*S <= resident warps), so that slice s of EVERY anchor starts at the same time; with row-sorted lists the
  // warps of a slice then walk the same band of bank rows together and each row comes from HBM once.
  const bool aligned = (variant & 0x100) != 0;
  const bool no_cta_fold = (variant & 0x80) != 0;   // bit 7: keep one slot per (warp, anchor) (A/B timing of the CTA-level fold)
  variant &= 0x1f;   // (bits 5 / 6 steer the sharded step's pre-pass and mean nothing here)
  if (!pick_variant(bank_dtype, (int)D, variant, &var))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_score: no kernel for this (dtype, feat_dim, variant); feat_dim in {32,64,128,256,512}");

  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  const int grid = di.sms * var.bps;
  long long NW = (long long)grid * kWarps;
  if (aligned && NW >= B) NW = (NW / B) * B;
  const long long P = B * K1;
  // CTA-level fold of the warp partials: possible when a CTA's range of pairs spans at most two anchors and its slots fit
  // static shared memory; not for the per-entry outputs mode (kept on the original layout) or the aligned partition
  const long long NC = filter_chunks(K1);
  const bool lists_ok = src != nullptr && src->compact != 0 && src->y != nullptr && full && !aligned && out_v1 == nullptr && NC >= 1 &&
                        B <= NW && row_end > row_begin;
  // band-sorted lists: wpu warps per anchor take interleaved blocks of the anchor's sorted unit lists
  long long wpu = 0;
  if (lists_ok && src->compact >= 3) wpu = banded_wpu(B, NC, NW, row_end - row_begin);
  const bool banded = wpu > 0;
  // (band sort asked for but not possible: the plain scan)
  const bool compact = lists_ok && !banded && src->compact < 3 && B * (NC + 1) < kMaxUnits;
  const bool cta_reduce = !compact && !banded && !aligned && D <= 256 && out_v1 == nullptr && (P * kWarps + NW - 1) / NW + 1 <= K1 && !no_cta_fold;
  const size_t need = workspace_bytes_for(B, K1, D, NW);
  if (workspace_bytes < need) return fail(CRDPN_E_WORKSPACE, "crdpn_crd_score: workspace too small");

  char* ws = (char*)workspace;
  unsigned int* ticket = (unsigned int*)ws;
  double* anchor_part = (double*)(ws + 16);
  float* slots = (float*)(ws + 16 + align_up((size_t)B * 8 * sizeof(double), 16));

  if (k_total < 0) return fail(CRDPN_E_BADARG, "crdpn_crd_score: k_total must be >= 0");
  const double Kd = (double)(k_total > 0 ? k_total : (K1 - 1));  // negatives per anchor over ALL shards
  const double Pn = 1.0 / (double)n_data;
  const float mPn_f = (float)(Kd * Pn);
  const float c_f = (float)(Kd * Pn + (double)eps);

  ScoreParams sp;
  sp.bank1 = (const char*)bank1;
  sp.bank2 = (const char*)bank2;
  sp.row_stride_bytes = (long long)row_stride * (long long)esz;
  sp.v1 = v1; sp.v2 = v2;
  sp.idx = (const long long*)contrast_idx;
  sp.idx_mode = src ? src->mode : 0;
  sp.idx32 = src ? src->idx32 : nullptr;
  sp.y = src ? (const long long*)src->y : nullptr;
  sp.seed = src ? src->seed : 0; sp.offset = src ? src->offset : 0;
  sp.offset_dev = (src && src->mode == 2) ? g_sampler_offset_dev : nullptr;
  sp.draw_n = src ? src->draw_n : 0; sp.draw_base = src ? src->draw_base : 0;
  sp.B = (int)B; sp.K1 = (int)K1; sp.D = (int)D;
  sp.row_begin = row_begin; sp.row_end = row_end;
  sp.k_exp = (float)(1.4426950408889634 / (double)T);
  sp.inv_Z1 = full ? (float)(1.0 / (double)Z1) : 0.f;
  sp.inv_Z2 = full ? (float)(1.0 / (double)Z2) : 0.f;
  sp.c = c_f;
  sp.inv_mPn = (Kd > 0) ? (float)(1.0 / (double)mPn_f) : 0.f;
  sp.eps_over_mPn = (Kd > 0) ? (float)(((double)c_f - (double)mPn_f) / (double)mPn_f) : 0.f;
  sp.inv_BT = (float)(1.0 / ((double)B * (double)T));
  sp.out_v1 = out_v1; sp.out_v2 = out_v2;
  sp.slots = slots;
  sp.maxseg = cta_reduce ? 2 : maxseg_for(P, NW, K1);
  sp.ticket = ticket;
  sp.nw = NW;
  sp.cta_reduce = cta_reduce ? 1 : 0;
  {
    static const int prefetch_on = [] { const char* e = getenv("CRDPN_SCORE_PREFETCH"); return (e && e[0] == '1') ? 1 : 0; }();
    sp.prefetch = prefetch_on;
  }
  sp.compact = banded ? 3 : compact ? 1 : 0;
  sp.wpu = (int)wpu;
  sp.band_mul = banded ? (unsigned)((((unsigned long long)kBands) << 32) / (unsigned long long)(row_end - row_begin)) + 1u : 0u;
  if (banded) sp.nw = B * wpu;
  sp.NC = (int)NC;
  char* cbase = ws + workspace_slots_end(B, K1, D, NW);
  long long* anchor_start = (long long*)cbase;
  int* ucount = (int*)(cbase + align_up((size_t)(B + 1) * 8, 256));
  int* cl = (int*)((char*)ucount + align_up((size_t)(B * NC) * 4, 256));
  sp.cl = cl; sp.ucount = ucount; sp.anchor_start = anchor_start;

  cudaStream_t st = (cudaStream_t)stream;
  if (compact || banded) {   // pre-pass: drop the entries other shards own, keep list order / sort by band (compact == 2, 4: the caller has launched it)
    if (sp.idx_mode != 2) sp.y = (const long long*)src->y;
    if (src->compact != 2 && src->compact != 4) {
      if (banded && !src->all_in_shard && (row_end - row_begin) * 2 <= n_data) crd_band_sort_kernel<true><<<(int)(B * NC), 256, 0, st>>>(sp, cl, ucount);
      else if (banded) crd_band_sort_kernel<false><<<(int)(B * NC), 256, 0, st>>>(sp, cl, ucount);
      else crd_shard_filter_kernel<<<(int)(B * NC), 256, 0, st>>>(sp, cl, ucount);
      CRDPN_LAUNCH_CHECK("crd_shard_filter_kernel");
    }
  }
  {
    ScopedKernelTimer tm(CRDPN_K_CRD_SCORE, st);
    (full ? var.full : var.sums)<<<grid, kThreads, 0, st>>>(sp);
  }
  CRDPN_LAUNCH_CHECK("crd_score_kernel");

  FinalizeParams fp;
  fp.slots = slots; fp.maxseg = sp.maxseg; fp.NW = banded ? sp.nw : NW; fp.group = cta_reduce ? kWarps : 1;
  fp.anchor_start = compact ? anchor_start : nullptr;
  fp.wpa = banded ? (int)wpu : 0;
  fp.B = (int)B; fp.K1 = (int)K1; fp.D = (int)D; fp.full = full ? 1 : 0;
  fp.grad_v1 = grad_v1; fp.grad_v2 = grad_v2;
  fp.anchor_part = anchor_part; fp.result = result; fp.ticket = ticket;
  fp.xchg = 0; fp.rank = 0; fp.world = 1; fp.reduced = nullptr;
  fp.off_ctl = fp.off_slots = fp.parity_stride = fp.slot_words = 0; fp.timeout = 0;
  for (int r = 0; r < p2p::kMaxWorld; ++r) fp.peers.buf[r] = nullptr;
  if (xchg != nullptr) {
    if (!full) return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: Z1 and Z2 must be frozen");
    fp.xchg = 1; fp.rank = xchg->rank; fp.world = xchg->world; fp.peers = xchg->peers; fp.reduced = xchg->reduced;
    fp.off_ctl = xchg->off_ctl; fp.off_slots = xchg->off_slots; fp.parity_stride = xchg->parity_stride;
    fp.slot_words = xchg->slot_words; fp.timeout = xchg->timeout;
  }
  if (upd != nullptr && upd->row_end > upd->row_begin) {
    const int ublocks = (int)((2 * B + (kFinalizeThreads / 32) - 1) / (kFinalizeThreads / 32));
    if (bank_dtype == CRDPN_F32)
      crd_finalize_update_kernel<float><<<(int)B + ublocks, kFinalizeThreads, 0, st>>>(fp, *upd);
    else
      crd_finalize_update_kernel<__nv_bfloat16><<<(int)B + ublocks, kFinalizeThreads, 0, st>>>(fp, *upd);
    CRDPN_LAUNCH_CHECK("crd_finalize_update_kernel");
  } else {
    crd_finalize_kernel<<<(int)B, kFinalizeThreads, 0, st>>>(fp);
    CRDPN_LAUNCH_CHECK("crd_finalize_kernel");
  }
  return CRDPN_OK;
}

extern "C" int crdpn_crd_score(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                               const float* v1, const float* v2, const int64_t* contrast_idx,
                               int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                               int64_t row_begin, int64_t row_end,
                               float T, float Z1, float Z2, float eps,
                               float* out_v1, float* out_v2, double* result, float* grad_v1, float* grad_v2,
                               void* workspace, size_t workspace_bytes, int variant, void* stream) {
  return score_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, contrast_idx, B, K1, D, n_data, k_total, row_begin, row_end,
                    T, Z1, Z2, eps, out_v1, out_v2, result, grad_v1, grad_v2, workspace, workspace_bytes, variant,
                    nullptr, stream);
}

// ---- bank-streaming step (crd_stream.cuh) ------------------------------------------------------------------------
struct TsLayout {
  size_t count, off, cursor, records, ahist, aoff, coarse_off, keys, partial, loss_part, total;
  int T, G, NB, GA, cshift;
  TsLayout(long long B, long long K1, long long rows, int sms, int tile_rows = ts::kTR) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    T = (int)((rows + tile_rows - 1) / tile_rows);
    G = sms;
    GA = sms * 2;
    // coarse buckets: pass B runs one CTA per bucket, so aim at >= ~0.8 CTAs per SM while the (CTA, bucket) table fits the scan
    cshift = ts::kCoarseShift;
    while (cshift > 5 && ((T + (1 << cshift) - 1) >> cshift) * 5 < sms * 4 &&
           (long long)((T + (1 << (cshift - 1)) - 1) >> (cshift - 1)) * GA <= ts::kPartScanMax)
      --cshift;
    NB = (T + (1 << cshift) - 1) >> cshift;   // pass-A CTAs: the two passes over idx are latency-bound per CTA, the scan is limited to 48 K (CTA, bucket) pairs
    size_t o = 0;
    count = o; o += up((size_t)(T + 1) * 4);
    off = o; o += up((size_t)(T + 1) * 4);
    cursor = o; o += up((size_t)(T + 1) * 4);
    records = o; o += up((size_t)(B * K1 + 8) * 4);
    ahist = o; o += up((size_t)GA * NB * 4);
    aoff = o; o += up((size_t)GA * NB * 4);
    coarse_off = o; o += up((size_t)(NB + 1) * 4);
    keys = o; o += up((size_t)(B * K1 + 8) * 4);
    partial = o; o += up((size_t)G * B * 2 * ts::kD * 4);
    loss_part = o; o += up((size_t)G * ts::kWarpsTS * 2 * 4);
    total = o;
  }
};

// can the streaming formulation run this problem?  (D = 128, B <= 48, interleaved or dense banks, tile table fits;
// fp32 banks: register kernel, 32-row tiles; bf16 banks: tensor-core kernel, 64-row tiles)
static bool ts_supported(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype, int64_t B, int64_t K1,
                         int64_t D, int64_t rows, int sms) {
  if ((bank_dtype != CRDPN_F32 && bank_dtype != CRDPN_BF16) || D != ts::kD || B < 1 || B > ts::kMaxB || rows < 1 ||
      B * K1 >= (1ll << 31))
    return false;
  const int64_t esz = bank_dtype == CRDPN_F32 ? 4 : 2;
  const bool inter = row_stride == 2 * D && (const char*)bank2 == (const char*)bank1 + D * esz;
  const bool dense = row_stride == D;
  if (!inter && !dense) return false;
  const long long tr = bank_dtype == CRDPN_F32 ? ts::kTR : tc::kRows;
  const long long T = (rows + tr - 1) / tr;
  return (T + sms - 1) / sms + 1 <= (bank_dtype == CRDPN_F32 ? ts::kMaxTilesPerCta : tc::kMaxTilesPerCta);
}

extern "C" int crdpn_crd_stream_workspace_bytes(int64_t B, int64_t K1, int64_t D, int64_t rows_local, int device, size_t* bytes) {
  if (!bytes || B <= 0 || K1 <= 0 || D <= 0 || rows_local <= 0) return fail(CRDPN_E_BADARG, "crdpn_crd_stream_workspace_bytes: bad argument");
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  *bytes = TsLayout(B, K1, rows_local, di.sms).total;
  return CRDPN_OK;
}

// SWIZZLE_128B tensor map over one bf16 bank [rows][128] with a row stride of row_stride elements, box {64, 64}
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_bank_tensor_map(CUtensorMap* tm, const void* bank, int64_t rows, int64_t row_stride) {
  static TensorMapEncodeTiledFn encode = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (TensorMapEncodeTiledFn)f;
  }();
  if (!encode) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step (streaming): cuTensorMapEncodeTiled is not available in this driver");
  const cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)tc::kRows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(bank), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step (streaming): cuTensorMapEncodeTiled rejected the bank layout");
  return CRDPN_OK;
}

static int stream_step_impl(void* bank1, void* bank2, int64_t row_stride, int bank_dtype, const float* v1, const float* v2,
                            const int64_t* contrast_idx, int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                            int64_t row_begin, int64_t row_end, float T, float Z1, float Z2, float eps, double* result,
                            float* grad_v1, float* grad_v2, void* workspace, size_t workspace_bytes, const UpdateParams& upd,
                            int copy_only, void* stream) {
  if (!v1 || !v2 || !contrast_idx || !result || !grad_v1 || !grad_v2 || !workspace || !bank1 || !bank2)
    return fail(CRDPN_E_BADARG, "crdpn_crd_step (streaming): null pointer");
  if (B <= 0 || K1 <= 0 || n_data <= 0 || row_end <= row_begin || !(T > 0.f) || k_total < 0)
    return fail(CRDPN_E_BADARG, "crdpn_crd_step (streaming): bad size");
  if (((uintptr_t)bank1 | (uintptr_t)bank2 | (uintptr_t)v1 | (uintptr_t)v2 | (uintptr_t)workspace) & 15)
    return fail(CRDPN_E_ALIGN, "crdpn_crd_step (streaming): 16-byte alignment required");
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  const int64_t rows = row_end - row_begin;
  if (!ts_supported(bank1, bank2, row_stride, bank_dtype, B, K1, D, rows, di.sms))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step (streaming): needs feat_dim 128, batch <= 48, interleaved or dense banks");
  const bool use_tc = bank_dtype == CRDPN_BF16;   // bf16 banks: tensor-core kernel (crd_tc_stream.cuh), 64-row tiles
  if (use_tc && copy_only) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step (streaming): copy-only probe is fp32 only");
  if (di.max_smem_optin < (use_tc ? (int)tc::kSmem : (int)ts::kSmemBytes))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step (streaming): not enough shared memory");
  const int tshift = use_tc ? 6 : 5;
  const TsLayout L(B, K1, rows, di.sms, 1 << tshift);
  if (workspace_bytes < L.total) return fail(CRDPN_E_WORKSPACE, "crdpn_crd_step (streaming): workspace too small (crdpn_crd_stream_workspace_bytes)");
  char* ws = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;

  ts::BucketParams bp;
  bp.idx = (const long long*)contrast_idx;
  bp.P = B * K1;
  bp.K1 = (unsigned)K1;
  bp.row_begin = row_begin; bp.row_end = row_end;
  bp.count = (unsigned*)(ws + L.count); bp.off = (unsigned*)(ws + L.off); bp.cursor = (unsigned*)(ws + L.cursor);
  bp.records = (unsigned*)(ws + L.records);
  bp.T = L.T;
  bp.vec_ok = ((uintptr_t)contrast_idx & 15) == 0 ? 1 : 0;
  bp.tshift = tshift;

  const double Kd = (double)(k_total > 0 ? k_total : (K1 - 1));
  const double Pn = 1.0 / (double)n_data;
  const float mPn_f = (float)(Kd * Pn);
  const float c_f = (float)(Kd * Pn + (double)eps);
  ts::StreamParams sp;
  sp.bank1 = (const char*)bank1; sp.bank2 = (const char*)bank2;
  sp.interleaved = row_stride == 2 * D ? 1 : 0;
  sp.rows = rows; sp.B = (int)B; sp.T = L.T;
  sp.v1 = v1; sp.v2 = v2;
  sp.tile_off = bp.off; sp.records = bp.records;
  sp.k_exp = (float)(1.4426950408889634 / (double)T);
  sp.inv_Z1 = (float)(1.0 / (double)Z1); sp.inv_Z2 = (float)(1.0 / (double)Z2);
  sp.c = c_f;
  sp.inv_mPn = (Kd > 0) ? (float)(1.0 / (double)mPn_f) : 0.f;
  sp.eps_over_mPn = (Kd > 0) ? (float)(((double)c_f - (double)mPn_f) / (double)mPn_f) : 0.f;
  sp.inv_BT = (float)(1.0 / ((double)B * (double)T));
  sp.partial = (float*)(ws + L.partial);
  sp.loss_part = (float*)(ws + L.loss_part);
  sp.copy_only = copy_only;

  static bool attr_set[64] = {false};
  if (!attr_set[device]) {
    CRDPN_CUDA(cudaFuncSetAttribute(ts::crd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::kSmemBytes));
    CRDPN_CUDA(cudaFuncSetAttribute(tc::crd_tc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmem));
    CRDPN_CUDA(cudaFuncSetAttribute(ts::ts_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::kScanSmemMax * 4));
    CRDPN_CUDA(cudaFuncSetAttribute(ts::ts_coarse_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ts::kPartScanMax * 4));
    attr_set[device] = true;
  }
  {
    ScopedKernelTimer tm(CRDPN_K_CRD_SCORE, st);   // the bucketing is part of what replaces the gather pass
    if ((long long)L.NB * L.GA <= ts::kPartScanMax && L.NB * 4 <= 48 * 1024) {
      // two-pass partition: shared-memory atomics only
      ts::PartParams pp;
      pp.idx = bp.idx; pp.P = bp.P; pp.K1 = bp.K1; pp.row_begin = row_begin; pp.row_end = row_end;
      pp.T = L.T; pp.NB = L.NB; pp.GA = L.GA; pp.tshift = tshift; pp.cshift = L.cshift;
      pp.ahist = (unsigned*)(ws + L.ahist); pp.aoff = (unsigned*)(ws + L.aoff); pp.coarse_off = (unsigned*)(ws + L.coarse_off);
      pp.keys = (unsigned*)(ws + L.keys); pp.tile_off = bp.off; pp.records = bp.records;
      const size_t sh = (size_t)L.NB * 4;
      ts::ts_coarse_hist_kernel<<<L.GA, ts::kPartThreads, sh, st>>>(pp);
      CRDPN_LAUNCH_CHECK("ts_coarse_hist_kernel");
      ts::ts_coarse_scan_kernel<<<1, 1024, (size_t)L.NB * L.GA * 4, st>>>(pp);
      CRDPN_LAUNCH_CHECK("ts_coarse_scan_kernel");
      ts::ts_coarse_scatter_kernel<<<L.GA, ts::kPartThreads, sh, st>>>(pp);
      CRDPN_LAUNCH_CHECK("ts_coarse_scatter_kernel");
      ts::ts_fine_kernel<<<L.NB, 512, 0, st>>>(pp);
      CRDPN_LAUNCH_CHECK("ts_fine_kernel");
    } else {  // very large shards: global-atomic counting sort
      CRDPN_CUDA(cudaMemsetAsync(bp.count, 0, (size_t)(L.T + 1) * 4, st));
      const int pre_grid = di.sms * 8;
      ts::ts_hist_kernel<<<pre_grid, 256, 0, st>>>(bp);
      CRDPN_LAUNCH_CHECK("ts_hist_kernel");
      ts::ts_scan_kernel<<<1, 1024, L.T <= ts::kScanSmemMax ? (size_t)L.T * 4 : 0, st>>>(bp);
      CRDPN_LAUNCH_CHECK("ts_scan_kernel");
      ts::ts_scatter_kernel<<<pre_grid, 256, 0, st>>>(bp);
      CRDPN_LAUNCH_CHECK("ts_scatter_kernel");
    }
    if (use_tc) {
      tc::TcParams tp;
      tp.bank1 = sp.bank1; tp.bank2 = sp.bank2; tp.interleaved = sp.interleaved; tp.rows = rows; tp.B = (int)B; tp.T = L.T;
      tp.v1 = v1; tp.v2 = v2; tp.tile_off = sp.tile_off; tp.records = sp.records;
      tp.k_exp = sp.k_exp; tp.inv_Z1 = sp.inv_Z1; tp.inv_Z2 = sp.inv_Z2; tp.c = sp.c; tp.inv_mPn = sp.inv_mPn;
      tp.eps_over_mPn = sp.eps_over_mPn; tp.inv_BT = sp.inv_BT; tp.partial = sp.partial; tp.loss_part = sp.loss_part;
      static const int tc_prof = getenv("CRDPN_TC_PROF") ? 1 : 0;
      tp.prof = tc_prof;
      CUtensorMap tm1, tm2;
      if ((rc = make_bank_tensor_map(&tm1, bank1, rows, row_stride)) != CRDPN_OK) return rc;
      if ((rc = make_bank_tensor_map(&tm2, bank2, rows, row_stride)) != CRDPN_OK) return rc;
      tc::crd_tc_stream_kernel<<<L.G, tc::kThreads, tc::kSmem, st>>>(tp, tm1, tm2);
      CRDPN_LAUNCH_CHECK("crd_tc_stream_kernel");
    } else {
      ts::crd_stream_kernel<<<L.G, ts::kThreadsTS, ts::kSmemBytes, st>>>(sp);
      CRDPN_LAUNCH_CHECK("crd_stream_kernel");
    }
  }
  ts::TsFinalizeParams fp;
  fp.partial = sp.partial; fp.loss_part = sp.loss_part; fp.tile_off = bp.off;
  fp.G = L.G; fp.B = (int)B; fp.T = L.T;
  fp.grad_v1 = grad_v1; fp.grad_v2 = grad_v2; fp.result = result;
  const int ublocks = (int)((2 * B + 7) / 8);
  if (use_tc) ts::ts_finalize_update_kernel<__nv_bfloat16><<<(int)B + ublocks, ts::kTsFinalizeThreads, 0, st>>>(fp, upd);
  else ts::ts_finalize_update_kernel<float><<<(int)B + ublocks, ts::kTsFinalizeThreads, 0, st>>>(fp, upd);
  CRDPN_LAUNCH_CHECK("ts_finalize_update_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_crd_step(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                              const float* v1, const float* v2, const int64_t* contrast_idx, const int64_t* y,
                              int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                              int64_t row_begin, int64_t row_end,
                              float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                              double* result, float* grad_v1, float* grad_v2,
                              void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!(Z1 > 0.f && Z2 > 0.f)) return fail(CRDPN_E_BADARG, "crdpn_crd_step: Z1 and Z2 must be frozen (> 0) before a full step");
  int rc = check_update_args(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end);
  if (rc) return rc;
  const UpdateParams u = make_update_params(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end,
                                            momentum, one_minus_momentum);
  if ((variant & 0x200) && (variant & 0x1000))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step: the bank-streaming kernels take an int64 contrast_idx");
  if (variant & 0x200)  // bank-streaming formulation (crd_stream.cuh); the workspace is crdpn_crd_stream_workspace_bytes
    return stream_step_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, contrast_idx, B, K1, D, n_data, k_total, row_begin,
                            row_end, T, Z1, Z2, eps, result, grad_v1, grad_v2, workspace, workspace_bytes, u,
                            (variant & 0x800) ? 1 : 0, stream);
  if (variant & 0x400) {    // band-sorted lists (crd_band_sort_kernel): repeats of a row within the step become L2 hits
    const IdxSource src{(variant & 0x1000) ? 1 : 0, (const int*)contrast_idx, y, 0, 0, 0, 0, 3};
    return score_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, (variant & 0x1000) ? nullptr : contrast_idx, B, K1, D, n_data,
                      k_total, row_begin, row_end, T, Z1, Z2, eps, nullptr, nullptr, result, grad_v1, grad_v2, workspace,
                      workspace_bytes, variant & ~0x1400, &u, stream, nullptr, &src);
  }
  if (variant & 0x1000) {   // contrast_idx is an int32 list
    const IdxSource src{1, (const int*)contrast_idx, nullptr, 0, 0, 0, 0, 0};
    return score_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, nullptr, B, K1, D, n_data, k_total, row_begin, row_end,
                      T, Z1, Z2, eps, nullptr, nullptr, result, grad_v1, grad_v2, workspace, workspace_bytes, variant & ~0x1000,
                      &u, stream, nullptr, &src);
  }
  return score_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, contrast_idx, B, K1, D, n_data, k_total, row_begin, row_end,
                    T, Z1, Z2, eps, nullptr, nullptr, result, grad_v1, grad_v2, workspace, workspace_bytes, variant,
                    &u, stream);
}

// crdpn_crd_step whose negatives are drawn INSIDE the scoring pass (uniform sampler): no index list exists in memory.
extern "C" int crdpn_crd_step_drawn(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                                    const float* v1, const float* v2, const int64_t* y,
                                    int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                                    int64_t row_begin, int64_t row_end,
                                    float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                                    uint64_t seed, uint64_t offset, int64_t draw_n, int64_t draw_base,
                                    double* result, float* grad_v1, float* grad_v2,
                                    void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!(Z1 > 0.f && Z2 > 0.f)) return fail(CRDPN_E_BADARG, "crdpn_crd_step_drawn: Z1 and Z2 must be frozen (> 0) before a full step");
  if (variant & 0x200) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step_drawn: the bank-streaming kernels need a materialised contrast_idx");
  if (draw_n <= 0 || draw_base < 0) return fail(CRDPN_E_BADARG, "crdpn_crd_step_drawn: bad sampler range");
  int rc = check_update_args(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end);
  if (rc) return rc;
  const UpdateParams u = make_update_params(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end,
                                            momentum, one_minus_momentum);
  const IdxSource src{2, nullptr, y, seed, offset, draw_n, draw_base, (variant & 0x400) ? 3 : 0};
  return score_impl(bank1, bank2, row_stride, bank_dtype, v1, v2, nullptr, B, K1, D, n_data, k_total, row_begin, row_end,
                    T, Z1, Z2, eps, nullptr, nullptr, result, grad_v1, grad_v2, workspace, workspace_bytes, variant & 0xfff,
                    &u, stream, nullptr, &src);
}

// The row-sharded step (one process per GPU) as one call: all-gather of the anchors over NVLink
// peer memory -> scoring pass over this rank's shard -> ONE kernel that reduces the per-warp partials, momentum-updates
// the positive rows this rank owns and sums the gradients / loss partials over the ranks (LL words pushed into every
// peer's slot, summed in rank order).  3 launches.  The bank-streaming variant keeps its own reduction kernel, so there
// the sum over ranks is the separate crdpn_p2p_allreduce_f32 launch.
// everything of the sharded step after the all-gather (shared by crdpn_crd_step_sharded and
// crdpn_crd_loss_forward_sharded, which may draw in-shard negatives in between)
int crdpn::sharded_step_core(void* bank1, void* bank2, int64_t row_stride, int bank_dtype, void* const* peer_bufs_host,
                             int rank, int world, int64_t Bmax, int64_t Dmax, const int64_t* contrast_idx, int64_t B,
                             int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
                             float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                             float* v1_all, float* v2_all, int64_t* y_all, float* partial, double* result, float* reduced,
                             void* workspace, size_t workspace_bytes, int variant, void* stream, int idx_mode, uint64_t seed,
                             uint64_t offset, int64_t draw_n, int64_t draw_base, const float* v1_local, const float* v2_local,
                             const int64_t* y_local, const int32_t* offs_host) {
  int rc;
  if ((variant & 0x200) && idx_mode != 0)
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_step_sharded: the bank-streaming kernels need a materialised int64 contrast_idx");
  // the pre-pass pays when most of the list belongs to other shards (measured on one rank's share of the headline step:
  // scoring pass 76 -> 67 us at 1/8 of the rows, nothing at 1/2); variant bit 6 forces it on (single-GPU tests), bit 5 off
  const bool gather = v1_local != nullptr;
  bool prefiltered = false;
  // bit 0x20: every entry of the list lives in this shard (in-shard negatives) -- no filter pre-pass; together with 0x400 it
  // still means band-sorted lists, sorted without the survivor compaction
  const bool all_in = (variant & 0x20) != 0 || (idx_mode == 2 && draw_base >= row_begin && draw_base + draw_n <= row_end);
  const bool want_compact = !(variant & 0x200) &&
                            ((variant & 0x400) || (!(variant & 0x20) && ((variant & 0x40) || (row_end - row_begin) * 3 <= n_data)));
  bool sweep = false;   // band-sorted lists (variant | 0x400)
  cudaStream_t st = (cudaStream_t)stream;
  SideStream* side = nullptr;
  if (want_compact && gather && workspace != nullptr) {
    // filter pre-pass on a second stream, beside the all-gather (it needs the index list only, not the anchors)
    int device = 0;
    CRDPN_CUDA(cudaGetDevice(&device));
    DeviceInfo di;
    if ((rc = device_info(device, &di)) != 0) return rc;
    CompactPlan cp;
    if (plan_compact(bank_dtype, D, variant & 0xf9f, B, K1, row_begin, row_end, di.sms, workspace, &cp) &&
        workspace_bytes >= workspace_bytes_for(B, K1, D, cp.NW) && (idx_mode != 0 || contrast_idx != nullptr)) {
      if ((rc = side_stream(&side)) != 0) return rc;
      ScoreParams fpar;
      memset(&fpar, 0, sizeof(fpar));
      fpar.idx = (const long long*)contrast_idx; fpar.idx32 = (const int*)contrast_idx; fpar.idx_mode = idx_mode;
      fpar.seed = seed; fpar.offset = offset; fpar.draw_n = draw_n; fpar.draw_base = draw_base;
      fpar.offset_dev = idx_mode == 2 ? g_sampler_offset_dev : nullptr;
      fpar.B = (int)B; fpar.K1 = (int)K1; fpar.D = (int)D; fpar.row_begin = row_begin; fpar.row_end = row_end; fpar.NC = (int)cp.NC;
      sweep = (variant & 0x400) && banded_wpu(B, cp.NC, cp.NW, row_end - row_begin) > 0;
      fpar.band_mul = sweep ? (unsigned)((((unsigned long long)kBands) << 32) / (unsigned long long)(row_end - row_begin)) + 1u : 0u;
      CRDPN_CUDA(cudaEventRecord(side->fork, st));
      CRDPN_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
      if (sweep && !all_in && (row_end - row_begin) * 2 <= n_data) crd_band_sort_kernel<true><<<(int)(B * cp.NC), 256, 0, side->s>>>(fpar, cp.cl, cp.ucount);
      else if (sweep) crd_band_sort_kernel<false><<<(int)(B * cp.NC), 256, 0, side->s>>>(fpar, cp.cl, cp.ucount);
      else crd_shard_filter_kernel<<<(int)(B * cp.NC), 256, 0, side->s>>>(fpar, cp.cl, cp.ucount);
      CRDPN_LAUNCH_CHECK("crd_shard_filter_kernel");
      CRDPN_CUDA(cudaEventRecord(side->join, side->s));
      prefiltered = true;
    }
  }
  if (gather) {
    rc = crdpn_p2p_allgather_anchors(v1_local, v2_local, y_local, D, offs_host, peer_bufs_host, rank, world, Bmax, Dmax, v1_all,
                                     v2_all, y_all, stream);
    if (rc) return rc;
  }
  if (prefiltered) CRDPN_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  if (variant & 0x200) {
    rc = crdpn_crd_step(bank1, bank2, row_stride, bank_dtype, v1_all, v2_all, contrast_idx, y_all, B, K1, D, n_data, k_total,
                        row_begin, row_end, T, Z1, Z2, eps, momentum, one_minus_momentum, result, partial, partial + B * D,
                        workspace, workspace_bytes, variant, stream);
    if (rc) return rc;
    return crdpn_p2p_allreduce_f32(partial, 2 * B * D, result, 8, reduced, peer_bufs_host, rank, world, Bmax, Dmax, stream);
  }
  rc = check_update_args(bank1, bank2, row_stride, bank_dtype, v1_all, v2_all, y_all, B, D, row_begin, row_end);
  if (rc) return rc;
  const UpdateParams u = make_update_params(bank1, bank2, row_stride, bank_dtype, v1_all, v2_all, y_all, B, D, row_begin,
                                            row_end, momentum, one_minus_momentum);
  const p2p::Layout L(Bmax, Dmax, world);
  FinalizeParams x;
  x.rank = rank; x.world = world; x.reduced = reduced;
  for (int r = 0; r < p2p::kMaxWorld; ++r) x.peers.buf[r] = r < world ? (char*)peer_bufs_host[r] : nullptr;
  for (int r = 0; r < world; ++r)
    if (!x.peers.buf[r]) return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: null peer buffer");
  x.off_ctl = L.ctl; x.off_slots = L.slots; x.parity_stride = L.parity_stride; x.slot_words = L.slot_words;
  x.timeout = p2p::poll_timeout_ticks();
  // an empty shard still takes part in the exchange: score_impl launches the reduction kernel either way
  const IdxSource src{idx_mode, (const int*)contrast_idx, y_all, seed, offset, draw_n, draw_base,
                      prefiltered ? (sweep ? 4 : 2) : (want_compact ? ((variant & 0x400) ? 3 : 1) : 0), all_in ? 1 : 0};
  return score_impl(bank1, bank2, row_stride, bank_dtype, v1_all, v2_all, idx_mode == 0 ? contrast_idx : nullptr, B, K1, D,
                    n_data, k_total, row_begin, row_end, T, Z1, Z2, eps, nullptr, nullptr, result, partial, partial + B * D,
                    workspace, workspace_bytes, variant & 0xf9f, &u, stream, &x, &src);
}

extern "C" int crdpn_crd_step_sharded(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                                      const float* v1_local, const float* v2_local, const int64_t* y_local,
                                      const int32_t* offs_host, void* const* peer_bufs_host, int rank, int world,
                                      int64_t Bmax, int64_t Dmax, const int64_t* contrast_idx,
                                      int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                                      int64_t row_begin, int64_t row_end,
                                      float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                                      float* v1_all, float* v2_all, int64_t* y_all, float* partial, double* result,
                                      float* reduced, void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!offs_host || !peer_bufs_host || !v1_all || !v2_all || !y_all || !partial || !result || !reduced)
    return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: null pointer");
  if (world < 1 || world > p2p::kMaxWorld || rank < 0 || rank >= world)
    return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: bad rank / world (world <= 8)");
  if (!(Z1 > 0.f && Z2 > 0.f)) return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: Z1 and Z2 must be frozen (> 0)");
  const int64_t B = offs_host[world];
  if (B <= 0 || B > Bmax || D > Dmax) return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: batch does not fit the exchange buffer");
  if (!v1_local || !v2_local || !y_local) return fail(CRDPN_E_BADARG, "crdpn_crd_step_sharded: null pointer");
  return sharded_step_core(bank1, bank2, row_stride, bank_dtype, peer_bufs_host, rank, world, Bmax, Dmax, contrast_idx, B, K1, D,
                           n_data, k_total, row_begin, row_end, T, Z1, Z2, eps, momentum, one_minus_momentum, v1_all, v2_all,
                           y_all, partial, result, reduced, workspace, workspace_bytes, variant & ~0x1000, stream,
                           (variant & 0x1000) ? 1 : 0, 0, 0, 0, 0, v1_local, v2_local, y_local, offs_host);
}

extern "C" int crdpn_crd_momentum_update(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                                         const float* v1, const float* v2, const int64_t* y,
                                         int64_t B, int64_t D, int64_t row_begin, int64_t row_end,
                                         float momentum, float one_minus_momentum, void* stream) {
  int rc = check_update_args(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end);
  if (rc) return rc;
  if (row_end == row_begin) return CRDPN_OK;
  const UpdateParams u = make_update_params(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, D, row_begin, row_end,
                                            momentum, one_minus_momentum);
  const int wpb = kFinalizeThreads / 32;
  const int grid = (int)((2 * B + wpb - 1) / wpb);
  cudaStream_t st = (cudaStream_t)stream;
  if (bank_dtype == CRDPN_F32) crd_update_kernel<float><<<grid, kFinalizeThreads, 0, st>>>(u);
  else crd_update_kernel<__nv_bfloat16><<<grid, kFinalizeThreads, 0, st>>>(u);
  CRDPN_LAUNCH_CHECK("crd_update_kernel");
  return CRDPN_OK;
}

static int alias_draw_impl(const float* prob, const int64_t* alias, int64_t n, int64_t count, uint64_t seed,
                           uint64_t offset, const int64_t* y, int64_t K1, int64_t row_base, int64_t* out, void* stream) {
  if ((prob == nullptr) != (alias == nullptr) || !out || n <= 0 || count < 0) return fail(CRDPN_E_BADARG, "crdpn_alias_draw: bad argument");
  if (count == 0) return CRDPN_OK;
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  long long blocks = (count + 255) / 256;
  const long long cap = (long long)di.sms * 8;
  if (blocks > cap) blocks = cap;
  alias_draw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(prob, (const long long*)alias, n, count, seed, offset,
                                                                  (const long long*)y, K1, row_base, (long long*)out);
  CRDPN_LAUNCH_CHECK("alias_draw_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_alias_draw(const float* prob, const int64_t* alias, int64_t n, int64_t count,
                                uint64_t seed, uint64_t offset, int64_t* out, void* stream) {
  return alias_draw_impl(prob, alias, n, count, seed, offset, nullptr, 1, 0, out, stream);
}

extern "C" int crdpn_alias_draw_contrast(const float* prob, const int64_t* alias, int64_t n, const int64_t* y,
                                         int64_t B, int64_t K1, uint64_t seed, uint64_t offset, int64_t* out,
                                         void* stream) {
  if (!y || B <= 0 || K1 <= 0) return fail(CRDPN_E_BADARG, "crdpn_alias_draw_contrast: bad argument");
  return alias_draw_impl(prob, alias, n, B * K1, seed, offset, y, K1, 0, out, stream);
}

extern "C" int crdpn_alias_draw_contrast_local(const float* prob, const int64_t* alias, int64_t n_local, int64_t row_base,
                                               const int64_t* y, int64_t B, int64_t K1, uint64_t seed, uint64_t offset,
                                               int64_t* out, void* stream) {
  if (!y || B <= 0 || K1 <= 0 || row_base < 0) return fail(CRDPN_E_BADARG, "crdpn_alias_draw_contrast_local: bad argument");
  return alias_draw_impl(prob, alias, n_local, B * K1, seed, offset, y, K1, row_base, out, stream);
}
