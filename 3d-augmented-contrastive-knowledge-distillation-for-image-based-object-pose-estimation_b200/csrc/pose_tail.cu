// pose_tail.cu -- the PoseEstimator tail (everything between the two encoders and the losses) as ONE persistent kernel.
//
// Reference: auxiliary/model.py:183-203 (DeformNet: four 1x1 Conv1d on a length-1 sequence = four Linear layers, three
// BatchNorm1d + ReLU, tanh), model.py:238-272 (cat of the shape and image features, six fc_* heads, the projector MLP);
// eval-mode call at KD/common/base_class.py:363, train mode at training.py:30,47.
//
// Shape of the work: a chain of skinny GEMMs, M = batch rows (<= 160: the reference trains with 160 and distils with
// 138 = 46 x 3 views) against 8.1 M weights.  2.2 GFLOP is nothing; the 32 MB of weights are: the chain is bound by
// streaming every weight once from HBM (~5 us) and by the five dependent layers.  So:
//   * every layer runs "swap-AB" on tcgen05: D[128 output channels x Npad batch rows] += W_tile[128 x 64 k] . X[Npad x 64 k]^T,
//     both operands K-major SWIZZLE_128B images that arrive by plain bulk copies (weights are packed into that image once,
//     activations are written in it by the producing layer's epilogue);
//   * fp32-accurate mode: every operand is a bf16 (hi, lo) pair, every product three MMAs (hi.hi + lo.hi + hi.lo), ~2^-17
//     relative per product, fp32 accumulation in TMEM -- the same bytes as fp32 weights, 1e-5 of the reference's outputs;
//     bf16 mode (north_star's 1e-2 tolerance mode) reads the hi planes only: half the bytes, one MMA per product;
//   * each layer is cut into (128-channel tile, K-split) tasks so that one dependency level fills the 148 SMs; a task's
//     partial accumulator goes to an L2-resident scratch, the splits of a tile meet on a counter, and each of them then
//     reduces 1/S of the tile's channels over the S partials in fixed order (deterministic), applies bias / BatchNorm /
//     activation and writes the next layer's operand image (and the fp32 outputs the caller asked for);
//   * layers wait for their source layer on device-side counters, not on kernel boundaries: the producer warp issues a
//     task's WEIGHT copies first and only then waits for the activations, so the weight stream of layer l+1 is already
//     in flight while layer l finishes.  One launch for cat + DeformNet + six heads + projector.
// Train mode (batch-statistics BatchNorm, training.py:30): the reduce step also has every batch row of its channels in
// hand, so mean / variance over the batch, the running-statistics update and the saved (mean, 1/std) cost no extra pass.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "pointnet_common.cuh"

namespace crdpn {
namespace pt {

using namespace pn;

constexpr int kEpi = 256;              // warps 0-7: epilogue (TMEM lane quadrant = warp % 4, column half = warp / 4) and reduce
constexpr int kPtThreads = kEpi + 64;  // warp 8: bulk-copy producer, warp 9: MMA issuer
constexpr int kMaxLayers = 10;
constexpr int kDxMaxSplits = 16;   // split-K ranges of the backward's dx GEMMs (crdpn_pose_tail_backward)
constexpr int kMaxTasks = 768;         // kernel-parameter space: 2 bytes per task
constexpr int kMaxGroups = 200;        // (layer, tile) pairs
constexpr uint32_t kWPlane = 16384;    // one [128 rows x 64 k] bf16 plane of a weight tile
constexpr int kCtrDone = 0;            // counters: [0] input images packed, [1 + l] tasks of layer l finished
constexpr int kCtrExit = 16;
constexpr int kCtrTile = 32;           // [32 + group]: splits of the tile whose partial is in the scratch
constexpr size_t kCtrBytes = 1024;
constexpr size_t kDbgBytes = (size_t)kMaxTasks * 8 * 8;   // optional per-task timestamps (CRDPN_POSE_TAIL_PROF), behind the counters

struct Layer {
  const char* w;        // [tiles][KB][2 planes][16 KB]
  const float* bias;    // [O]
  const char* x;        // source image [KB][2 planes][Npad x 128 B]
  char* y;              // image of this layer's output (null: nobody consumes it)
  float* out;           // [B][O] fp32 (null: not wanted)
  float* scratch;       // [tiles][S][Npad][128] fp32 partial accumulators
  const float* gamma;   // train-mode BatchNorm (null: none)
  const float* beta;
  float* run_mean;
  float* run_var;
  float* save_mean;
  float* save_istd;
  float* xhat;          // train mode: [B][O] normalised activations (z - mean) / std, for backward
  int KB, tiles, O, act, dep, S, ykb, first_group;
};

struct Input {          // an input image: concat(a [B, Fa], b [B, Fb]) along k, zero padded to KB * 64
  const float* a;
  const float* b;
  char* img;
  int Fa, Fb, KB;
};

struct Params {
  Layer L[kMaxLayers];
  Input in[2];
  unsigned* ctr;
  int nlayers, ntasks, B, Npad, NS, planes, train;
  uint32_t stg_bytes;          // shared-memory staging area of the split-K reduce
  float bn_momentum, bn_eps;
  long long timeout;
  unsigned long long* dbg;     // null, or [ntasks][8] globaltimer ns: accumulator ready, TMEM drained, partial stored (fenced), tile
                               // complete, reduce loop left, after __threadfence, after the proxy fence, task done
  unsigned short tasks[kMaxTasks];   // layer | tile << 4 | split << 11, in dependency order
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// spin until *p >= target (another CTA's release); traps after `timeout` clocks: a protocol bug, never a legitimate wait
__device__ __forceinline__ void spin_ge(const unsigned* p, unsigned target, long long timeout) {
  const long long t0 = clock64();
  while (ld_acquire(p) < target) {
    if (clock64() - t0 > timeout) {
      printf("pose_tail_kernel: CTA %d gave up waiting for counter %p (%u < %u)\n", (int)blockIdx.x, (const void*)p, ld_acquire(p), target);
      __trap();
    }
  }
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
// fp32 -> (hi, lo) bf16 pair: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  lo = pack_bf16(a - bf16_lo(hi), b - bf16_hi(hi));
}

// ------------------------------------------------------------------------------------------------------------------
// weights [O, I] fp32 row-major (optionally scaled per row: eval-mode BatchNorm folded) -> (hi, lo) operand image
__global__ void __launch_bounds__(256) pose_tail_pack_kernel(const float* __restrict__ W, const float* __restrict__ row_scale,
                                                             int O, int I, int KB, int tiles, char* __restrict__ image) {
  const long long chunks = (long long)tiles * 128 * KB * 8;   // 8 k per chunk
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % (KB * 8)), o = (int)(i / (KB * 8));
    const int k0 = kc * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (o < O) {
      const float s = row_scale ? row_scale[o] : 1.f;
      if ((I & 3) == 0 && k0 + 8 <= I) {
        const float4 a = *reinterpret_cast<const float4*>(W + (size_t)o * I + k0);
        const float4 b = *reinterpret_cast<const float4*>(W + (size_t)o * I + k0 + 4);
        v[0] = a.x * s; v[1] = a.y * s; v[2] = a.z * s; v[3] = a.w * s;
        v[4] = b.x * s; v[5] = b.y * s; v[6] = b.z * s; v[7] = b.w * s;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k0 + j < I) v[j] = W[(size_t)o * I + k0 + j] * s;
      }
    }
    uint4 hi, lo;
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
    const int tile = o >> 7, kb = k0 >> 6;
    char* base = image + ((size_t)(tile * KB + kb) * 2) * kWPlane + sw128_off(o & 127, k0 & 63);
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + kWPlane) = lo;
  }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPtThreads, 1) pose_tail_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t xplane = (uint32_t)p.Npad * 128u;
  const uint32_t stage_bytes = 2 * kWPlane + 2 * xplane;
  const uint32_t off_bar = (uint32_t)p.NS * stage_bytes;
  const uint32_t bar_full = base + off_bar, bar_empty = bar_full + 64, bar_accf = bar_full + 128, bar_acce = bar_full + 136,
                 bar_stg = bar_full + 152;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + off_bar + 144);
  float* zstat = reinterpret_cast<float*>(sm + off_bar + 256);          // train mode: [128][3] mean, 1/std, gamma
  float* par_s = zstat + 128 * 3;                                        // [3][128]: bias, gamma, beta of the current tile
  float4* stg = reinterpret_cast<float4*>(sm + off_bar + 256 + 128 * 6 * 4);   // [S][rows x groups] partial slices
  const uint32_t stg_addr = base + off_bar + 256 + 128 * 6 * 4;

  if (tid == 0) {
    for (int i = 0; i < p.NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 1);
    mbar_init(bar_stg, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256u);

  // ---- input images: concat + (hi, lo) split of the fp32 features, rows >= B and columns >= F zero -------------------
  for (int q = 0; q < 2; ++q) {
    const Input& in = p.in[q];
    if (in.img == nullptr) continue;
    const int chunks = p.Npad * in.KB * 8;
    const bool vec = (in.Fa & 7) == 0 && (in.Fb & 7) == 0;
    for (int i = blockIdx.x * kPtThreads + tid; i < chunks; i += gridDim.x * kPtThreads) {
      const int n = i / (in.KB * 8), k0 = (i - n * in.KB * 8) * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (n < p.B) {
        if (vec) {
          const float* src = k0 < in.Fa ? in.a + (size_t)n * in.Fa + k0 : (k0 < in.Fa + in.Fb ? in.b + (size_t)n * in.Fb + (k0 - in.Fa) : nullptr);
          if (src) {
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            if (k < in.Fa) v[j] = in.a[(size_t)n * in.Fa + k];
            else if (k < in.Fa + in.Fb) v[j] = in.b[(size_t)n * in.Fb + (k - in.Fa)];
          }
        }
      }
      uint4 hi, lo;
      split2(v[0], v[1], hi.x, lo.x);
      split2(v[2], v[3], hi.y, lo.y);
      split2(v[4], v[5], hi.z, lo.z);
      split2(v[6], v[7], hi.w, lo.w);
      char* dst = in.img + (size_t)(k0 >> 6) * 2 * xplane + sw128_off(n, k0 & 63);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + xplane) = lo;
    }
  }
  __threadfence();
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) atomicAdd(p.ctr + kCtrDone, 1u);
  const uint32_t tmem = *tmem_slot;

  auto decode = [&](int t, int& l, int& tile, int& s) {
    const unsigned e = p.tasks[t];
    l = (int)(e & 0xfu); tile = (int)((e >> 4) & 0x7fu); s = (int)(e >> 11);
  };

  if (warp == kEpi / 32) {
    // ================================ producer: one elected lane issues every bulk copy =============================
    if (elect_one()) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.ntasks; t += gridDim.x) {
        int l, tile, s;
        decode(t, l, tile, s);
        const Layer& L = p.L[l];
        const int kb0 = L.KB * s / L.S, kb1 = L.KB * (s + 1) / L.S;
        bool gated = false;
        for (int kb = kb0; kb < kb1; ++kb, ++g) {
          const uint32_t st = g % (uint32_t)p.NS, ph = (g / (uint32_t)p.NS) & 1u;
          mbar_wait(bar_empty + 8 * st, ph ^ 1u);
          const uint32_t dst = base + st * stage_bytes, bar = bar_full + 8 * st;
          mbar_expect_tx(bar, (uint32_t)p.planes * (kWPlane + xplane));
          bulk_g2s(dst, L.w + ((size_t)(tile * L.KB + kb) * 2) * kWPlane, (uint32_t)p.planes * kWPlane, bar);
          if (!gated) {   // the weights are on their way; the activations need the source layer to have finished
            const unsigned target = L.dep < 0 ? gridDim.x : (unsigned)(p.L[L.dep].tiles * p.L[L.dep].S);
            spin_ge(p.ctr + kCtrDone + 1 + L.dep, target, p.timeout);
            fence_proxy_async_all();
            gated = true;
          }
          bulk_g2s(dst + 2 * kWPlane, L.x + (size_t)kb * 2 * xplane, (uint32_t)p.planes * xplane, bar);
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpi / 32 + 1) {
    // ================================ MMA issuer ====================================================================
    const uint32_t idesc = make_idesc(128u, (uint32_t)p.Npad);
    uint32_t g = 0, done_tasks = 0;
    for (int t = blockIdx.x; t < p.ntasks; t += gridDim.x, ++done_tasks) {
      int l, tile, s;
      decode(t, l, tile, s);
      const Layer& L = p.L[l];
      const int nkb = L.KB * (s + 1) / L.S - L.KB * s / L.S;
      mbar_wait(bar_acce, (done_tasks & 1u) ^ 1u);   // the epilogue has drained the previous task's accumulator
      tc_fence_after();
      for (int i = 0; i < nkb; ++i, ++g) {
        const uint32_t st = g % (uint32_t)p.NS, ph = (g / (uint32_t)p.NS) & 1u;
        mbar_wait(bar_full + 8 * st, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t w_hi = base + st * stage_bytes, x_hi = w_hi + 2 * kWPlane;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = umma_desc_sw128(w_hi + ks * 32), b_hi = umma_desc_sw128(x_hi + ks * 32);
            umma_f16(tmem, a_hi, b_hi, idesc, (i > 0 || ks > 0) ? 1u : 0u);
            if (p.planes == 2) {
              umma_f16(tmem, umma_desc_sw128(w_hi + kWPlane + ks * 32), b_hi, idesc, 1u);
              umma_f16(tmem, a_hi, umma_desc_sw128(x_hi + xplane + ks * 32), idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * st);
          if (i == nkb - 1) umma_commit(bar_accf);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue warps (TMEM lane quadrant = warp) ====================================
    uint32_t done_tasks = 0, stg_phase = 0;
    const int Npad = p.Npad;
    for (int t = blockIdx.x; t < p.ntasks; t += gridDim.x, ++done_tasks) {
      int l, tile, s;
      decode(t, l, tile, s);
      const Layer& L = p.L[l];
      const int S = L.S;
      float* tile_scr = L.scratch + (size_t)tile * S * Npad * 128;
      // ---- this split's partial accumulator: TMEM -> scratch [n][128 channels] (lane = channel: coalesced)
      mbar_wait(bar_accf, done_tasks & 1u);
      tc_fence_after();
      if (p.dbg != nullptr && tid == 0) p.dbg[8 * t] = gtime();
      {
        const int q = warp & 3, half = warp >> 2;
        const int cmid = ((Npad >> 4) + 1) / 2 * 16, cbeg = half ? cmid : 0, cend = half ? Npad : cmid;
        float* part = tile_scr + (size_t)s * Npad * 128 + q * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        int c0 = cbeg;
        for (; c0 + 64 <= cend; c0 += 64) {   // two TMEM loads in flight per wait
          uint32_t v[32], w[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) __stcg(part + (size_t)(c0 + i) * 128, __uint_as_float(v[i]));
#pragma unroll
          for (int i = 0; i < 32; ++i) __stcg(part + (size_t)(c0 + 32 + i) * 128, __uint_as_float(w[i]));
        }
        for (; c0 + 32 <= cend; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) __stcg(part + (size_t)(c0 + i) * 128, __uint_as_float(v[i]));
        }
        if (c0 < cend) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) __stcg(part + (size_t)(c0 + i) * 128, __uint_as_float(v[i]));
        }
      }
      if (p.dbg != nullptr && tid == 0) p.dbg[8 * t + 1] = gtime();
      if (tid < 128) {   // the tile's bias / BatchNorm affine: one global round trip here instead of one per reduce iteration
        const int c = tile * 128 + tid;
        const bool real = c < L.O;
        par_s[tid] = real ? L.bias[c] : 0.f;
        par_s[128 + tid] = (real && L.gamma != nullptr) ? L.gamma[c] : 0.f;
        par_s[256 + tid] = (real && L.gamma != nullptr) ? L.beta[c] : 0.f;
      }
      tc_fence_before();
      named_sync(1, kEpi);
      unsigned* tile_ctr = p.ctr + kCtrTile + L.first_group + tile;
      if (tid == 0) {
        mbar_arrive(bar_acce);
        __threadfence();   // cumulative: releases the whole CTA's partial stores (ordered before this by the barrier)
        if (p.dbg != nullptr) p.dbg[8 * t + 2] = gtime();
        atomicAdd(tile_ctr, 1u);
        spin_ge(tile_ctr, (unsigned)S, p.timeout);
        if (p.dbg != nullptr) p.dbg[8 * t + 3] = gtime();
      }
      named_sync(1, kEpi);
      // ---- reduce 1/S of the tile's channels over the S partials (fixed order), bias / BatchNorm / activation, outputs
      const int g0 = 32 * s / S, g1 = 32 * (s + 1) / S;   // 4-channel groups of the tile
      const bool bn = p.train != 0 && L.gamma != nullptr;
      const int act = L.act, O = L.O, ycols = L.ykb * 64, Breal = p.B;
      float* const outp = L.out;
      char* const yimg = L.y;
      auto emit = [&](int n, int c, float (&v)[4]) {   // activation, fp32 output, (hi, lo) image of the next layer's operand
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (act == 1) v[j] = fmaxf(v[j], 0.f);
          else if (act == 2) v[j] = tanhf(v[j]);
        }
        if (outp != nullptr && n < Breal && c < O) {
          if ((O & 3) == 0) {
            *reinterpret_cast<float4*>(outp + (size_t)n * O + c) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c + j < O) outp[(size_t)n * O + c + j] = v[j];
          }
        }
        if (yimg != nullptr && c < ycols) {
          uint2 hi, lo;
          split2(v[0], v[1], hi.x, lo.x);
          split2(v[2], v[3], hi.y, lo.y);
          char* dst = yimg + (size_t)(c >> 6) * 2 * xplane + sw128_off(n, c & 63);
          *reinterpret_cast<uint2*>(dst) = hi;
          *reinterpret_cast<uint2*>(dst + xplane) = lo;
        }
      };
      if (!bn) {
        // This split's share of the tile = a range of batch ROWS (all 128 channels): in the [n][128] scratch that is one
        // contiguous run per split, so S bulk copies bring all S partial slices into shared memory at once (one L2 round
        // trip whatever S is); (row, 4-channel group) pairs are then summed in split order (deterministic).
        const int ra = Npad * s / S, rb = Npad * (s + 1) / S;
        const int rpp = max(1, min(rb - ra, (int)(p.stg_bytes / (uint32_t)(S * 512))));
        for (int r0 = ra; r0 < rb; r0 += rpp) {
          const int rows = min(rpp, rb - r0), pairs = rows * 32;
          if (tid == 0) {
            fence_proxy_async_all();   // the other splits' generic stores (acquired above) -> this async-proxy read
            mbar_expect_tx(bar_stg, (uint32_t)(S * rows) * 512u);
            for (int s2 = 0; s2 < S; ++s2)
              bulk_g2s(stg_addr + (uint32_t)(s2 * rows) * 512u, tile_scr + ((size_t)s2 * Npad + r0) * 128, (uint32_t)rows * 512u, bar_stg);
          }
          mbar_wait(bar_stg, stg_phase);
          stg_phase ^= 1u;
          if (p.dbg != nullptr && tid == 0 && r0 == ra) p.dbg[8 * t + 5] = gtime();
          for (int i = tid; i < pairs; i += kEpi) {
            float4 a = stg[i];
            for (int s2 = 1; s2 < S; ++s2) {
              const float4 b = stg[s2 * pairs + i];
              a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const int grp = i & 31;
            float v[4] = {a.x + par_s[grp * 4], a.y + par_s[grp * 4 + 1], a.z + par_s[grp * 4 + 2], a.w + par_s[grp * 4 + 3]};
            emit(r0 + (i >> 5), tile * 128 + grp * 4, v);
          }
          named_sync(1, kEpi);   // the next pass (or the next task) overwrites the staging area
        }
      } else {
        // batch-statistics BatchNorm needs every batch row of a channel in one place: the share is a range of 4-channel
        // GROUPS (all rows), staged by 16-byte asynchronous copies, as many groups per pass as the staging area holds
        const int gpp = max(1, (int)(p.stg_bytes / (uint32_t)(S * Npad * 16)));
        for (int gc = g0; gc < g1; gc += gpp) {
          const int ngrp = min(gpp, g1 - gc), pairs = Npad * ngrp;
          for (int e = tid; e < pairs * S; e += kEpi) {
            const int s2 = e / pairs, i = e - s2 * pairs;
            const int nr = i / ngrp, grp = gc + (i - nr * ngrp);
            const float* src = tile_scr + ((size_t)s2 * Npad + nr) * 128 + grp * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg_addr + (uint32_t)e * 16u), "l"(src) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          named_sync(1, kEpi);
          for (int i = tid; i < pairs; i += kEpi) {
            float4 a = stg[i];
            for (int s2 = 1; s2 < S; ++s2) {
              const float4 b = stg[s2 * pairs + i];
              a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            stg[i] = a;   // (slot 0 of this pair: only this thread touches it)
          }
          named_sync(1, kEpi);
          // one warp per channel: mean, then the centred second moment (two passes over shared memory: no cancellation)
          const float* zs = reinterpret_cast<const float*>(stg);
          for (int ch = warp; ch < ngrp * 4; ch += kEpi / 32) {
            const float* z = zs + ch;   // row n of channel ch: z[n * ngrp * 4]
            float sum = 0.f;
            for (int n = lane; n < p.B; n += 32) sum += z[(size_t)n * ngrp * 4];
            sum = warp_sum_xor(sum, 16, 1);
            const float mean = sum / (float)p.B;
            float sq = 0.f;
            for (int n = lane; n < p.B; n += 32) { const float d = z[(size_t)n * ngrp * 4] - mean; sq = fmaf(d, d, sq); }
            sq = warp_sum_xor(sq, 16, 1);
            const float var = sq / (float)p.B;
            const int ct = gc * 4 + ch, c = tile * 128 + ct;   // channel inside the tile / of the layer
            const float istd = rsqrtf(var + p.bn_eps);
            if (lane == 0) {
              zstat[3 * ch] = mean;
              zstat[3 * ch + 1] = istd;
              zstat[3 * ch + 2] = par_s[128 + ct];
              if (c < L.O) {
                // the Linear / Conv bias cancels inside train-mode BatchNorm; it only enters the running mean
                L.save_mean[c] = mean + par_s[ct];
                L.save_istd[c] = istd;
                if (L.run_mean != nullptr) {
                  L.run_mean[c] = (1.f - p.bn_momentum) * L.run_mean[c] + p.bn_momentum * (mean + par_s[ct]);
                  const float unbiased = p.B > 1 ? var * (float)p.B / (float)(p.B - 1) : var;
                  L.run_var[c] = (1.f - p.bn_momentum) * L.run_var[c] + p.bn_momentum * unbiased;
                }
              }
            }
          }
          named_sync(1, kEpi);
          for (int i = tid; i < pairs; i += kEpi) {
            const int nr = i / ngrp, gl = i - nr * ngrp;
            const int c = tile * 128 + (gc + gl) * 4;
            const float4 a = stg[i];
            const float zv[4] = {a.x, a.y, a.z, a.w};
            float v[4], xh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ch = gl * 4 + j;
              xh[j] = (zv[j] - zstat[3 * ch]) * zstat[3 * ch + 1];
              v[j] = fmaf(xh[j], zstat[3 * ch + 2], par_s[256 + (gc + gl) * 4 + j]);
            }
            if (L.xhat != nullptr && nr < p.B && c < L.O) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j < L.O) L.xhat[(size_t)nr * L.O + c + j] = xh[j];
            }
            emit(nr, c, v);
          }
          named_sync(1, kEpi);   // the next pass (or the next task) overwrites the staging area
        }
      }
      if (p.dbg != nullptr && tid == 0) p.dbg[8 * t + 4] = gtime();
      named_sync(1, kEpi);
      if (tid == 0) {
        __threadfence();
        fence_proxy_async_all();
        atomicAdd(p.ctr + kCtrDone + 1 + l, 1u);
        if (p.dbg != nullptr) p.dbg[8 * t + 7] = gtime();
      }
    }
  }

  // ---- teardown: TMEM back, and the LAST CTA out resets every counter (each CTA is past all of its waits by now)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256u);
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(p.ctr + kCtrExit, 1u) == gridDim.x - 1) {
      for (int i = 0; i < (int)(kCtrBytes / 4); ++i) p.ctr[i] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct Plan {
  int n = 0, B = 0, Npad = 0, NS = 0;
  int KB[kMaxLayers], tiles[kMaxLayers], S[kMaxLayers], ykb[kMaxLayers], level[kMaxLayers], first_group[kMaxLayers];
  int in_kb[2] = {0, 0};
  size_t off_in[2], off_y[kMaxLayers], off_scr[kMaxLayers], total = 0;
  std::vector<unsigned short> tasks;
  size_t smem = 0, stg = 0;
};

static int make_plan(const crdpn_pose_tail_layer* layers, int n, int64_t B, int64_t Fs, int64_t Fi, int sms, int smem_optin, bool train,
                     Plan* pl) {
  if (!layers || n < 1 || n > kMaxLayers) return fail(CRDPN_E_BADARG, "pose tail: 1..10 layers");
  if (B < 1 || B > 256) return fail(CRDPN_E_UNSUPPORTED, "pose tail: batch rows must be 1..256 per call");
  if (Fs < 0 || Fi < 1) return fail(CRDPN_E_BADARG, "pose tail: bad feature dimensions");
  pl->n = n;
  pl->B = (int)B;
  pl->Npad = (int)((B + 15) / 16 * 16);
  pl->in_kb[0] = (int)((Fs + Fi + 63) / 64);
  pl->in_kb[1] = (int)((Fi + 63) / 64);
  int groups = 0, maxlevel = 0;
  for (int l = 0; l < n; ++l) {
    const auto& a = layers[l];
    if (!a.weights || !a.bias || a.O < 1 || a.I < 1 || a.src < -2 || a.src >= l || a.act < 0 || a.act > 2)
      return fail(CRDPN_E_BADARG, "pose tail: bad layer description");
    const int64_t src_width = a.src == -1 ? Fs + Fi : a.src == -2 ? Fi : layers[a.src].O;
    if (a.I != src_width) return fail(CRDPN_E_BADARG, "pose tail: layer input width does not match its source");
    pl->KB[l] = (int)((a.I + 63) / 64);
    pl->tiles[l] = (int)((a.O + 127) / 128);
    pl->ykb[l] = 0;
    pl->level[l] = a.src < 0 ? 1 : pl->level[a.src] + 1;
    maxlevel = std::max(maxlevel, pl->level[l]);
    pl->first_group[l] = groups;
    groups += pl->tiles[l];
    if (pl->tiles[l] > 127) return fail(CRDPN_E_UNSUPPORTED, "pose tail: layer wider than 16256 outputs");
  }
  if (groups > kMaxGroups) return fail(CRDPN_E_UNSUPPORTED, "pose tail: too many output tiles");
  for (int l = 0; l < n; ++l)
    if (layers[l].src >= 0) pl->ykb[layers[l].src] = std::max(pl->ykb[layers[l].src], pl->KB[l]);
  const size_t xplane = (size_t)pl->Npad * 128;
  const size_t stage = 2 * kWPlane + 2 * xplane;
  const size_t fixed = 256 + 128 * 6 * 4 + 1024;               // barriers, BatchNorm statistics + tile parameters, alignment slack
  const size_t want_stg = (size_t)32 * pl->Npad * 16;          // one pass when S divides 32 (every share is 32 / S groups)
  if ((size_t)smem_optin < fixed + 2 * stage + 16384) return fail(CRDPN_E_UNSUPPORTED, "pose tail: batch too large for two pipeline stages");
  pl->NS = 2;
  while (pl->NS < 4 && fixed + (size_t)(pl->NS + 1) * stage + want_stg <= (size_t)smem_optin) ++pl->NS;
  pl->stg = std::min(want_stg, ((size_t)smem_optin - fixed - (size_t)pl->NS * stage) / 16 * 16);
  pl->smem = (size_t)pl->NS * stage + fixed + pl->stg;
  // a BatchNorm share stages whole channel groups (every batch row): at least one group x S splits has to fit
  const int max_split = (int)std::max<size_t>(1, std::min<size_t>(32, pl->stg / ((size_t)pl->Npad * 16)));
  (void)train;
  // K-splits: per dependency level, the smallest K-blocks-per-task u for which the level's tasks fit one wave of CTAs
  pl->tasks.clear();
  for (int lev = 1; lev <= maxlevel; ++lev) {
    int u = 1;
    for (;; ++u) {
      long long cnt = 0;
      for (int l = 0; l < n; ++l)
        if (pl->level[l] == lev) cnt += (long long)pl->tiles[l] * std::min(32, (pl->KB[l] + u - 1) / u);
      if (cnt <= sms || u >= 4096) break;
    }
    for (int l = 0; l < n; ++l) {
      if (pl->level[l] != lev) continue;
      pl->S[l] = std::max(1, std::min(std::min(32, pl->KB[l]), (pl->KB[l] + u - 1) / u));
      pl->S[l] = std::max(1, std::min(pl->S[l], max_split));
      for (int tile = 0; tile < pl->tiles[l]; ++tile)
        for (int s = 0; s < pl->S[l]; ++s) pl->tasks.push_back((unsigned short)((unsigned)l | ((unsigned)tile << 4) | ((unsigned)s << 11)));
    }
  }
  if ((int)pl->tasks.size() > kMaxTasks) return fail(CRDPN_E_UNSUPPORTED, "pose tail: too many tasks");
  // workspace: counters | input images | layer images | partial scratch
  size_t off = kCtrBytes + kDbgBytes;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  for (int q = 0; q < 2; ++q) pl->off_in[q] = take((size_t)pl->in_kb[q] * 2 * xplane);
  for (int l = 0; l < n; ++l) pl->off_y[l] = take((size_t)pl->ykb[l] * 2 * xplane);
  for (int l = 0; l < n; ++l) pl->off_scr[l] = take((size_t)pl->tiles[l] * pl->S[l] * pl->Npad * 128 * 4);
  pl->total = off;
  return CRDPN_OK;
}

}  // namespace pt
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_pose_tail_image_bytes(int64_t O, int64_t I, size_t* bytes) {
  if (O < 1 || I < 1 || !bytes) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_image_bytes: bad argument");
  *bytes = (size_t)((O + 127) / 128) * (size_t)((I + 63) / 64) * 2 * pt::kWPlane;
  return CRDPN_OK;
}

extern "C" int crdpn_pose_tail_pack_weights(const float* W, const float* row_scale, int64_t O, int64_t I, void* image, void* stream) {
  if (!W || !image || O < 1 || I < 1) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_pack_weights: bad argument");
  if ((reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(image) & 1023))
    return fail(CRDPN_E_ALIGN, "crdpn_pose_tail_pack_weights: W must be 16-byte, image 1024-byte aligned");
  const int KB = (int)((I + 63) / 64), tiles = (int)((O + 127) / 128);
  const long long chunks = (long long)tiles * 128 * KB * 8;
  const int grid = (int)std::min<long long>((chunks + 255) / 256, 148 * 8);
  pt::pose_tail_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, row_scale, (int)O, (int)I, KB, tiles, (char*)image);
  CRDPN_LAUNCH_CHECK("pose_tail_pack_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_pose_tail_workspace_bytes(const crdpn_pose_tail_layer* layers, int n_layers, int64_t B, int64_t shape_dim,
                                               int64_t img_dim, size_t* bytes) {
  if (!bytes) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_workspace_bytes: bad argument");
  DeviceInfo di;
  if (int rc = device_info(-1, &di)) return rc;
  pt::Plan pl;
  if (int rc = pt::make_plan(layers, n_layers, B, shape_dim, img_dim, di.sms, di.max_smem_optin, false, &pl)) return rc;
  *bytes = pl.total;
  return CRDPN_OK;
}

extern "C" int crdpn_pose_tail_forward(const crdpn_pose_tail_layer* layers, int n_layers, const float* shape_feature,
                                       const float* img_feature, int64_t B, int64_t shape_dim, int64_t img_dim, int flags,
                                       float bn_momentum, float bn_eps, void* workspace, size_t workspace_bytes, void* stream) {
  if (!img_feature || !workspace || (shape_dim > 0 && !shape_feature))
    return fail(CRDPN_E_BADARG, "crdpn_pose_tail_forward: null pointer");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(CRDPN_E_ALIGN, "crdpn_pose_tail_forward: workspace must be 1024-byte aligned");
  DeviceInfo di;
  if (int rc = device_info(-1, &di)) return rc;
  const bool train = (flags & CRDPN_POSE_TAIL_TRAIN) != 0;
  pt::Plan pl;
  if (int rc = pt::make_plan(layers, n_layers, B, shape_dim, img_dim, di.sms, di.max_smem_optin, train, &pl)) return rc;
  if (workspace_bytes < pl.total) return fail(CRDPN_E_WORKSPACE, "crdpn_pose_tail_forward: workspace too small");
  char* ws = (char*)workspace;
  pt::Params p;
  memset(&p, 0, sizeof(p));
  p.ctr = reinterpret_cast<unsigned*>(ws);
  p.nlayers = n_layers;
  p.ntasks = (int)pl.tasks.size();
  p.B = pl.B;
  p.Npad = pl.Npad;
  p.NS = pl.NS;
  p.stg_bytes = (uint32_t)pl.stg;
  p.planes = (flags & CRDPN_POSE_TAIL_BF16) ? 1 : 2;
  p.train = train ? 1 : 0;
  p.bn_momentum = bn_momentum;
  p.bn_eps = bn_eps;
  {
    // how long a CTA may wait for another CTA's counter before it gives up (clock ticks; ~2 GHz).  Generous by default: the
    // kernel is launched cooperatively, so every CTA is resident and a wait only ever lasts microseconds
    static const long long ticks = [] {
      const char* e = getenv("CRDPN_POSE_TAIL_TIMEOUT_S");
      const double sec = e ? atof(e) : 30.0;
      return (long long)((sec > 0.01 ? sec : 0.01) * 2.0e9);
    }();
    p.timeout = ticks;
  }
  p.dbg = (flags & CRDPN_POSE_TAIL_PROF) ? reinterpret_cast<unsigned long long*>(ws + pt::kCtrBytes) : nullptr;
  p.in[0] = pt::Input{shape_feature, img_feature, ws + pl.off_in[0], (int)shape_dim, (int)img_dim, pl.in_kb[0]};
  p.in[1] = pt::Input{img_feature, nullptr, ws + pl.off_in[1], (int)img_dim, 0, pl.in_kb[1]};
  bool use_in[2] = {false, false};
  for (int l = 0; l < n_layers; ++l) {
    const auto& a = layers[l];
    pt::Layer& L = p.L[l];
    if (reinterpret_cast<uintptr_t>(a.weights) & 1023) return fail(CRDPN_E_ALIGN, "crdpn_pose_tail_forward: weight image must be 1024-byte aligned");
    if (a.out && (reinterpret_cast<uintptr_t>(a.out) & 15)) return fail(CRDPN_E_ALIGN, "crdpn_pose_tail_forward: outputs must be 16-byte aligned");
    if (train && a.gamma && !(a.beta && a.save_mean && a.save_istd)) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_forward: train-mode BatchNorm needs beta, save_mean, save_istd");
    L.w = (const char*)a.weights;
    L.bias = a.bias;
    L.x = a.src < 0 ? ws + pl.off_in[-1 - a.src] : ws + pl.off_y[a.src];
    if (a.src < 0) use_in[-1 - a.src] = true;
    L.y = pl.ykb[l] > 0 ? ws + pl.off_y[l] : nullptr;
    L.out = a.out;
    L.scratch = reinterpret_cast<float*>(ws + pl.off_scr[l]);
    L.gamma = train ? a.gamma : nullptr;
    L.beta = a.beta; L.run_mean = a.running_mean; L.run_var = a.running_var; L.save_mean = a.save_mean; L.save_istd = a.save_istd; L.xhat = train ? a.xhat : nullptr;
    L.KB = pl.KB[l]; L.tiles = pl.tiles[l]; L.O = (int)a.O; L.act = a.act; L.dep = a.src < 0 ? -1 : a.src; L.S = pl.S[l];
    L.ykb = pl.ykb[l]; L.first_group = pl.first_group[l];
  }
  for (int q = 0; q < 2; ++q)
    if (!use_in[q]) p.in[q].img = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  memcpy(p.tasks, pl.tasks.data(), pl.tasks.size() * sizeof(unsigned short));
  static bool attr_set[64] = {false};
  int dev = 0;
  CRDPN_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    CRDPN_CUDA(cudaFuncSetAttribute(pt::pose_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
    attr_set[dev & 63] = true;
  }
  // cooperative launch: the CTAs wait for each other on device-side counters, so all of them have to be resident -- the
  // runtime guarantees it (or refuses the launch) instead of this code assuming an idle GPU
  void* kargs[] = {(void*)&p};
  CRDPN_CUDA(cudaLaunchCooperativeKernel((const void*)pt::pose_tail_kernel, dim3(di.sms), dim3(pt::kPtThreads), kargs, pl.smem, st));
  CRDPN_LAUNCH_CHECK("pose_tail_kernel");
  return CRDPN_OK;
}

// =====================================================================================================================
// Train-mode backward of the chain (training.py:75 differentiates model.py:183-203, 238-272 through autograd) as ONE call:
// per layer, last to first,
//   pull-back   g_a = g_y * act'(y);  BatchNorm (batch statistics): d_beta = sum_n g_a, d_gamma = sum_n g_a * xhat,
//               g_z = gamma / std * (g_a - d_beta / B - xhat * d_gamma / B);  d_bias = sum_n g_z        (one kernel)
//   dW = g_z^T x_in   (fp32 FFMA GEMM, K = batch rows; the concat input is two column ranges of dW)      (1-2 launches)
//   dx = g_z W        accumulated into the gradient of the layer's source (a layer output, or the two inputs)
// Everything is fp32 on CUDA cores: 4.4 GFLOP per step, ~0.2 ms -- the forward's split-precision tensor-core path is not
// needed for parity here and the GEMMs are K = 138..2048 against <= 2048 x 2048 outputs.
namespace crdpn {
namespace pt {

// thread (c, r): channel c of a 32-channel group, rows r, r + 8, ...; partial sums meet in shared memory
__global__ void __launch_bounds__(256) pose_tail_pullback_kernel(const float* __restrict__ g_ext, const float* __restrict__ g_acc,
                                                                 const float* __restrict__ y, const float* __restrict__ xhat,
                                                                 const float* __restrict__ gamma, const float* __restrict__ istd,
                                                                 int act, int B, int O, float* __restrict__ gz,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 float* __restrict__ dbias) {
  __shared__ float s_a[8][33], s_b[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool live = c < O;
  const bool bn = gamma != nullptr;
  float sb = 0.f, sg = 0.f;
  if (live) {
    for (int n = ry; n < B; n += 8) {
      const size_t i = (size_t)n * O + c;
      float ga = (g_ext ? g_ext[i] : 0.f) + (g_acc ? g_acc[i] : 0.f);
      const float yv = y[i];
      if (act == 1) ga = yv > 0.f ? ga : 0.f;
      else if (act == 2) ga *= 1.f - yv * yv;
      gz[i] = ga;
      sb += ga;
      if (bn) sg = fmaf(ga, xhat[i], sg);
    }
  }
  s_a[ry][cx] = sb;
  s_b[ry][cx] = sg;
  __syncthreads();
  float tb = 0.f, tg = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) { tb += s_a[r][cx]; tg += s_b[r][cx]; }   // fixed order: deterministic
  if (!bn) {
    if (live && ry == 0) dbias[c] = tb;
    return;
  }
  float sz = 0.f;
  if (live) {
    const float k = gamma[c] * istd[c], mb = tb / (float)B, mg = tg / (float)B;
    for (int n = ry; n < B; n += 8) {
      const size_t i = (size_t)n * O + c;
      const float v = k * (gz[i] - mb - xhat[i] * mg);
      gz[i] = v;
      sz += v;
    }
  }
  __syncthreads();
  s_a[ry][cx] = sz;
  __syncthreads();
  if (live && ry == 0) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += s_a[r][cx];
    dbias[c] = t;       // a bias in front of batch-statistics BatchNorm: zero up to rounding
    dgamma[c] = tg;
    dbeta[c] = tb;
  }
}

// C[m, n] (+)= sum_k A(m, k) * Bm[k * ldb + n];  A(m, k) = A[k * lda + m] (AK: "k-major", the g_z^T x case) or A[m * lda + k].
// 16 x 16 threads, each RM x 4 outputs of a (16 RM) x 64 tile; k in chunks of 16 through shared memory; fp32 FFMA.
// Split-K: blockIdx.z takes the k range [z * kspan, min(K, (z + 1) * kspan)) (kspan a multiple of 16) and writes its partial
// product to C + z * zstride; pose_tail_splitk_sum_kernel adds the partials in z order (deterministic).  gridDim.z == 1 with
// kspan >= K is the plain GEMM.
template <int RM, bool AK>
__global__ void __launch_bounds__(256) pose_tail_sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                                              float* __restrict__ C, int ldc, int M, int N, int Ktot, int accumulate,
                                                              int kspan, size_t zstride) {
  constexpr int TM = 16 * RM;
  __shared__ float As[16][TM + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * 64;
  const int kbeg = (int)blockIdx.z * kspan;
  const int K = min(Ktot, kbeg + kspan);   // end of this CTA's k range
  C += (size_t)blockIdx.z * zstride;
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < K; k0 += 16) {
#pragma unroll
    for (int e = t; e < 16 * TM; e += 256) {
      int k, m;
      if (AK) { k = e / TM; m = e - k * TM; } else { m = e >> 4; k = e & 15; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = AK ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int e = t; e < 16 * 64; e += 256) {
      const int k = e >> 6, n = e & 63;
      const int gk = k0 + k, gn = n0 + n;
      Bs[k][n] = (gk < K && gn < N) ? Bm[(size_t)gk * ldb + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int gm = m0 + ty * RM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) {
        float* dst = C + (size_t)gm * ldc + gn;
        *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
      }
    }
  }
}

// dst[m, n] (+)= sum_z P[z][m][n], z = 0 .. S-1 in that order;  P: S dense [M, N] partials, dst: pitch lddst.  N % 4 == 0 and
// 16-byte aligned rows take the float4 path.
__global__ void __launch_bounds__(256) pose_tail_splitk_sum_kernel(const float* __restrict__ P, int S, size_t zstride, float* __restrict__ dst,
                                                                   int lddst, int M, int N, int accumulate) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float a = accumulate ? dst[(size_t)m * lddst + n] : 0.f;
    for (int z = 0; z < S; ++z) a += P[(size_t)z * zstride + i];
    dst[(size_t)m * lddst + n] = a;
  }
}

// dW[O, I-range] = g_z^T x   (A = g_z [B, O] read k-major, B = x [B, ldx])
static int launch_dw(const float* gz, int O, const float* x, int ldx, int ncols, float* dW, int ldw, int B, cudaStream_t st) {
  dim3 grid((ncols + 63) / 64, (O + 63) / 64);
  pose_tail_sgemm_kernel<4, true><<<grid, 256, 0, st>>>(gz, O, x, ldx, dW, ldw, O, ncols, B, 0, (B + 15) / 16 * 16, 0);
  CRDPN_LAUNCH_CHECK("pose_tail_sgemm_kernel<dW>");
  return CRDPN_OK;
}
// how many k ranges the dx GEMM of a [B, ncols] result over K = O is cut into: its (ncols / 64) x (B / 32) tiles alone leave
// most SMs idle and make every CTA walk O / 16 dependent load -> sync -> FMA rounds (128 at O = 2048); aim at ~1000 CTAs of
// >= 4 rounds each
static int dx_splits(int O, int ncols, int B) {
  const int tiles = ((ncols + 63) / 64) * ((B + 31) / 32);
  int S = (1000 + tiles - 1) / tiles;
  const int smax = (O + 63) / 64;
  if (S > smax) S = smax;
  if (S > kDxMaxSplits) S = kDxMaxSplits;
  return S < 1 ? 1 : S;
}
// dx[B, ncols] (+)= g_z W[:, col0 : col0 + ncols]   (A = g_z [B, O] row-major, B = W [O, ldw]); `part`: kDxMaxSplits * B * ncols floats
static int launch_dx(const float* gz, int O, const float* W, int ldw, int ncols, float* dx, int lddx, int B, int accumulate, float* part,
                     cudaStream_t st) {
  const int S = dx_splits(O, ncols, B);
  if (S <= 1) {
    dim3 grid((ncols + 63) / 64, (B + 31) / 32);
    pose_tail_sgemm_kernel<2, false><<<grid, 256, 0, st>>>(gz, O, W, ldw, dx, lddx, B, ncols, O, accumulate, (O + 15) / 16 * 16, 0);
    CRDPN_LAUNCH_CHECK("pose_tail_sgemm_kernel<dx>");
    return CRDPN_OK;
  }
  const int kspan = ((O + S - 1) / S + 15) / 16 * 16;
  const int Sz = (O + kspan - 1) / kspan;          // ranges that are not empty (<= S)
  const size_t zstride = (size_t)B * ncols;
  dim3 grid((ncols + 63) / 64, (B + 31) / 32, Sz);
  pose_tail_sgemm_kernel<2, false><<<grid, 256, 0, st>>>(gz, O, W, ldw, part, ncols, B, ncols, O, 0, kspan, zstride);
  CRDPN_LAUNCH_CHECK("pose_tail_sgemm_kernel<dx, split-K>");
  const size_t total = zstride;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  pose_tail_splitk_sum_kernel<<<blocks, 256, 0, st>>>(part, Sz, zstride, dx, lddx, B, ncols, accumulate);
  CRDPN_LAUNCH_CHECK("pose_tail_splitk_sum_kernel");
  return CRDPN_OK;
}

}  // namespace pt
}  // namespace crdpn

extern "C" int crdpn_pose_tail_backward_workspace_bytes(const crdpn_pose_tail_bwd_layer* layers, int n_layers, int64_t B, size_t* bytes) {
  if (!layers || !bytes || n_layers < 1 || n_layers > pt::kMaxLayers || B < 1) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward_workspace_bytes: bad argument");
  size_t tot = 0, widest = 0;
  for (int l = 0; l < n_layers; ++l) {
    tot += 2 * (((size_t)B * (size_t)layers[l].O * 4 + 255) / 256 * 256);
    if ((size_t)layers[l].I > widest) widest = (size_t)layers[l].I;
  }
  tot += (size_t)pt::kDxMaxSplits * (size_t)B * widest * 4;   // split-K partials of the widest dx GEMM
  *bytes = tot;
  return CRDPN_OK;
}

extern "C" int crdpn_pose_tail_backward(const crdpn_pose_tail_bwd_layer* layers, int n_layers, const float* shape_feature,
                                        const float* img_feature, int64_t B, int64_t shape_dim, int64_t img_dim,
                                        float* d_shape_feature, float* d_img_feature, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (!layers || n_layers < 1 || n_layers > pt::kMaxLayers || !img_feature || !workspace || B < 1 || (shape_dim > 0 && !shape_feature))
    return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward: bad argument");
  size_t need = 0;
  if (int rc = crdpn_pose_tail_backward_workspace_bytes(layers, n_layers, B, &need)) return rc;
  if (workspace_bytes < need) return fail(CRDPN_E_WORKSPACE, "crdpn_pose_tail_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* gacc[pt::kMaxLayers];
  float* gz[pt::kMaxLayers];
  bool have[pt::kMaxLayers];
  char* ws = (char*)workspace;
  for (int l = 0; l < n_layers; ++l) {
    const auto& a = layers[l];
    if (!a.W || !a.y || !a.dW || !a.db || a.O < 1 || a.I < 1 || a.src < -2 || a.src >= l || (a.gamma && !(a.xhat && a.istd && a.dgamma && a.dbeta)))
      return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward: bad layer description");
    const size_t sz = ((size_t)B * (size_t)a.O * 4 + 255) / 256 * 256;
    gacc[l] = reinterpret_cast<float*>(ws); ws += sz;
    gz[l] = reinterpret_cast<float*>(ws); ws += sz;
    have[l] = false;
  }
  float* part = reinterpret_cast<float*>(ws);   // behind the per-layer areas: kDxMaxSplits * B * max(I) floats
  bool have_sf = false, have_img = false;
  const int Bi = (int)B, Fs = (int)shape_dim, Fi = (int)img_dim;
  for (int l = n_layers - 1; l >= 0; --l) {
    const auto& a = layers[l];
    const int O = (int)a.O, I = (int)a.I;
    if (!a.g_out && !have[l]) {   // nothing flows into this layer's output: all of its gradients are zero
      CRDPN_CUDA(cudaMemsetAsync(a.dW, 0, (size_t)O * I * 4, st));
      CRDPN_CUDA(cudaMemsetAsync(a.db, 0, (size_t)O * 4, st));
      if (a.gamma) {
        CRDPN_CUDA(cudaMemsetAsync(a.dgamma, 0, (size_t)O * 4, st));
        CRDPN_CUDA(cudaMemsetAsync(a.dbeta, 0, (size_t)O * 4, st));
      }
      continue;
    }
    pt::pose_tail_pullback_kernel<<<(O + 31) / 32, 256, 0, st>>>(a.g_out, have[l] ? gacc[l] : nullptr, a.y, a.xhat, a.gamma, a.istd,
                                                               a.act, Bi, O, gz[l], a.dgamma, a.dbeta, a.db);
    CRDPN_LAUNCH_CHECK("pose_tail_pullback_kernel");
    int rc = 0;
    if (a.src == -1) {
      if (I != Fs + Fi) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward: concat layer width");
      if (Fs > 0 && (rc = pt::launch_dw(gz[l], O, shape_feature, Fs, Fs, a.dW, I, Bi, st))) return rc;
      if ((rc = pt::launch_dw(gz[l], O, img_feature, Fi, Fi, a.dW + Fs, I, Bi, st))) return rc;
      if (d_shape_feature && Fs > 0) {
        if ((rc = pt::launch_dx(gz[l], O, a.W, I, Fs, d_shape_feature, Fs, Bi, have_sf ? 1 : 0, part, st))) return rc;
        have_sf = true;
      }
      if (d_img_feature) {
        if ((rc = pt::launch_dx(gz[l], O, a.W + Fs, I, Fi, d_img_feature, Fi, Bi, have_img ? 1 : 0, part, st))) return rc;
        have_img = true;
      }
    } else if (a.src == -2) {
      if (I != Fi) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward: image layer width");
      if ((rc = pt::launch_dw(gz[l], O, img_feature, Fi, Fi, a.dW, I, Bi, st))) return rc;
      if (d_img_feature) {
        if ((rc = pt::launch_dx(gz[l], O, a.W, I, Fi, d_img_feature, Fi, Bi, have_img ? 1 : 0, part, st))) return rc;
        have_img = true;
      }
    } else {
      const auto& s = layers[a.src];
      if (I != (int)s.O) return fail(CRDPN_E_BADARG, "crdpn_pose_tail_backward: layer input width does not match its source");
      if ((rc = pt::launch_dw(gz[l], O, s.y, I, I, a.dW, I, Bi, st))) return rc;
      if ((rc = pt::launch_dx(gz[l], O, a.W, I, I, gacc[a.src], I, Bi, have[a.src] ? 1 : 0, part, st))) return rc;
      have[a.src] = true;
    }
  }
  if (d_shape_feature && !have_sf && Fs > 0) CRDPN_CUDA(cudaMemsetAsync(d_shape_feature, 0, (size_t)B * Fs * 4, st));
  if (d_img_feature && !have_img) CRDPN_CUDA(cudaMemsetAsync(d_img_feature, 0, (size_t)B * Fi * 4, st));
  return CRDPN_OK;
}
