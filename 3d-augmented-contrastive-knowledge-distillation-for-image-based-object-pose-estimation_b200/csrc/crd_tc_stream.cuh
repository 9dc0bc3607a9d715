// crd_tc_stream.cuh -- bank-streaming CRD step on TENSOR CORES, first version: bf16 banks (included by crd_kernels.cu).
//
// The register formulation of the streaming step (crd_stream.cuh) reads every resident row once but is instruction-bound:
// ~140 CUDA-core instructions per sample for the two dot products and the two gradient updates.  Here that arithmetic
// runs on tcgen05: per 64-row tile (both banks stacked as ONE operand, M = 128)
//   scores     S[128 x 96]   = rows (K-major)  . [V2 | V1]^T          8 MMAs (M 128, N 96, K 16)
//   gradients  G2^T[128 x 48] += bank1 rows^T (the same image, MN-major) . C2     4 MMAs (M 128 features, N 48, K 16 rows)
//              G1^T[128 x 48] += bank2 rows^T . C1                                  4 MMAs
// where C1 / C2 [row][anchor] are the sparse coefficient matrices dL/ds the sample stage builds from the tile's records:
// one thread per record looks its two scores up in the shared-memory dump of S, adds the loss terms, and COUNTS itself on
// its (anchor, row) slot with a native integer atomic.  Records that repeat a slot (sampling with replacement) share the
// score, so the slot's coefficient is count x d: after a barrier the slot's first record stores that product, rounded to
// bf16 once, into the coefficient operand image (no floating-point atomics anywhere), and clears it again when the
// gradient MMAs have read it.  The gradient accumulators stay in TMEM for the whole kernel.  Tiles arrive by TMA: four
// SWIZZLE_128B tensor-map boxes per tile land as the operand image (4-stage ring, one mbarrier per stage); with 16-bit
// operands the one image serves both GEMMs (csrc/umma_tf32_probe.cu, mode 2).
// Roles: warps 0..7 dump and sample; warp 8 issues the TMA loads and the MMAs.  Iteration `it`, two __syncthreads per tile:
//   (a) dump S(it) TMEM -> smem                                  | #1 |
//   (b) workers: count + loss of tile it   ;  MMA warp: gradient MMAs of tile it-1, score MMAs of tile it+1      | #2 |
//       ... then refills the stage of tile it-1 with tile it+3
//   (c) workers: slot owners store coefficients of tile it; clear the coefficients of tile it-1
// Tiles with more records than kRecRegs = 4 per worker thread (1024 per tile: small banks, large K) take a generic
// three-barrier path.  Measured phase costs and what bounds the kernel today: DESIGN.md sections 4.13 and 8
// (CRDPN_TC_PROF=1 prints CTA 0's per-phase cycle counts).
// Restrictions: bf16 banks (north_star's 1e-2 tolerance mode: the anchors' embeddings and the coefficients are rounded to
// bf16 for the MMAs too), D = 128, B <= 48, interleaved or dense banks, step mode only.  fp32 banks need TF32 with
// round-to-nearest staging and a second (BASE32B) image for the MN-major operand: next round.
#pragma once

namespace tc {

using namespace pn;   // tcgen05 / mbarrier helpers of pointnet_common.cuh

constexpr int kRows = 64;                       // bank rows per tile
constexpr int kWorkers = 256;                   // warps 0..7: loads, score dump, sample stage
constexpr int kThreads = kWorkers + 32;         // warp 8: issues the MMAs
constexpr int kStages = 4;
constexpr uint32_t kAStage = 32768;             // 2 K-blocks x [128 stacked rows x 128 B]
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffBV = kStages * kAStage;                  // [V2 | V1]: 2 K-blocks x [96 x 128 B] = 24576
constexpr uint32_t kCImg = 6144;                                // one coefficient image: [48 anchors x 64 rows] bf16
constexpr uint32_t kOffC = kOffBV + 24576;                      // [tile parity][C2^T | C1^T]
constexpr uint32_t kSdPitch = 49;                               // floats per row of the score dump (odd: conflict-free)
constexpr uint32_t kOffSd = kOffC + 4 * kCImg;                  // scores [2 banks][64 rows][49] fp32 = 25088
constexpr uint32_t kOffCnt = kOffSd + 2 * 64 * kSdPitch * 4;    // records per (anchor, row) slot of the current tile: [48][64] u32
constexpr uint32_t kPosOne = 1u << 28;                          // ... negatives counted in bits [0, 28), the positive above
constexpr int kMaxTilesPerCta = 2047;
constexpr uint32_t kOffTab = kOffCnt + 48 * 64 * 4;             // tile_off slice of this CTA
constexpr uint32_t kOffBar = kOffTab + (kMaxTilesPerCta + 1) * 4;
constexpr uint32_t kSmem = kOffBar + 64 + 1024;
constexpr int kRecRegs = 4;                     // records per thread per tile handled on the fast path (1024 per tile)

struct TcParams {
  const char* bank1;
  const char* bank2;
  int interleaved;          // 1: [rows][2][128] bf16 in one allocation; 0: two dense [rows][128] arrays
  long long rows;
  int B, T;
  const float* v1;
  const float* v2;
  const unsigned* tile_off;
  const unsigned* records;
  float k_exp, inv_Z1, inv_Z2, c, inv_mPn, eps_over_mPn, inv_BT;
  float* partial;           // [grid][B][256]
  float* loss_part;         // [grid][16][2]
  int prof;                 // 1: CTA 0 prints its per-phase cycle counts (CRDPN_TC_PROF=1; debugging aid)
};

// mbarrier wait whose common case (the phase has already completed) costs one try_wait: the clocked, trapping loop of
// pn::mbar_wait only starts when the first probe fails
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (!done) mbar_wait(bar, parity);
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) {   // named barrier: ids 1 (all 9 warps) and 2 (the 8 worker warps)
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// one SWIZZLE_128B box {64 bf16, 64 rows} of a bank -> 8 KB of the operand image; completes on the stage's mbarrier
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
// shared-memory descriptor, SWIZZLE_128B (layout type 2)
__device__ __forceinline__ uint64_t desc128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// byte offset (inside one coefficient image) of the bf16 holding (anchor b, bank row r)
__device__ __forceinline__ uint32_t coef_off(unsigned b, unsigned r) { return sw128_off((int)b, (int)r); }

__global__ void __launch_bounds__(kThreads, 1) crd_tc_stream_kernel(const TcParams p, const __grid_constant__ CUtensorMap tm1,
                                                                    const __grid_constant__ CUtensorMap tm2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar_s = base + kOffBar, bar_g0 = base + kOffBar + 8, bar_g1 = base + kOffBar + 16;
  const uint32_t bar_full = base + kOffBar + 24;   // [kStages]: the tile's four TMA boxes have landed
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kOffBar + 56);
  unsigned* offs = reinterpret_cast<unsigned*>(sm + kOffTab);
  unsigned* cnt = reinterpret_cast<unsigned*>(sm + kOffCnt);
  float* Sd = reinterpret_cast<float*>(sm + kOffSd);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < 8;

  const int t_begin = (int)((long long)p.T * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.T * (blockIdx.x + 1) / gridDim.x);
  const int ntiles = t_end - t_begin;
  for (int i = tid; i <= ntiles; i += kThreads) offs[i] = p.tile_off[t_begin + i];
  // [V2 | V1] operand image (bf16, anchors >= B are zero rows), zeroed coefficient images and slot counters
  for (int i = tid; i < 96 * 32; i += kThreads) {
    const int n = i >> 5, e = (i & 31) * 4;
    const int b = n < 48 ? n : n - 48;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < p.B) v = *reinterpret_cast<const float4*>((n < 48 ? p.v2 : p.v1) + (size_t)b * 128 + e);
    *reinterpret_cast<uint2*>(sm + kOffBV + (e >> 6) * (96 * 128) + sw128_off(n, e & 63)) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  for (int i = tid; i < (int)(4 * kCImg / 16); i += kThreads) reinterpret_cast<uint4*>(sm + kOffC)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 48 * 64; i += kThreads) cnt[i] = 0u;
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_g0, 1);
    mbar_init(bar_g1, 1);
    for (int i = 0; i < kStages; ++i) mbar_init(bar_full + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto active = [&](int j) { return j >= 0 && j < ntiles && offs[j + 1] != offs[j]; };   // somebody sampled tile j
  // one tile = 64 rows x (256 B of bank 1 | 256 B of bank 2): four TMA boxes of 64 bf16 x 64 rows, each landing as one
  // 8 KB (bank, K-block) quarter of the stacked K-major SWIZZLE_128B operand image (row m = bank * 64 + r); rows past the
  // end of the shard are zero-filled by the tensor map (0 * C stays 0 in the gradient GEMM).  Every tile of the CTA is
  // loaded, sampled or not, so that stage j % kStages sees exactly one barrier phase per tile.  One thread (MMA warp).
  auto load_tile = [&](int j) {
    const uint32_t dst = base + kOffA + (uint32_t)(j % kStages) * kAStage, bar = bar_full + 8 * (uint32_t)(j % kStages);
    const int y = (t_begin + j) * kRows;
    mbar_expect_tx(bar, kAStage);
    tma_box(dst, &tm1, 0, y, bar);
    tma_box(dst + 8192, &tm2, 0, y, bar);
    tma_box(dst + 16384, &tm1, 64, y, bar);
    tma_box(dst + 16384 + 8192, &tm2, 64, y, bar);
  };
  constexpr uint32_t kIS = idesc_bf16(128, 96, 0), kIG = idesc_bf16(128, 48, 1);
  // Descriptors: ONE base per operand image, the per-k-step operands are the base plus a compile-time constant in the
  // 14-bit address field (shared memory is < 256 KB, so the sum never carries into the LBO field).  Building each of the
  // 32 descriptors of a tile from its address (and, shift, or, pack: a dependent chain in the one issuing thread) was most of
  // the ~82 cycles a tcgen05.mma cost that thread.
  const uint64_t bv_desc = desc128(base + kOffBV, 16, 1024);                 // [V2 | V1], K-major, resident
  const uint64_t c_desc0 = desc128(base + kOffC, 16, 1024);                  // coefficient images of tile parity 0 (C2^T)
  const uint64_t c_desc1 = desc128(base + kOffC + 2 * kCImg, 16, 1024);      // ... parity 1
  bool g_started = false;   // (MMA warp) the gradient accumulators hold something
  auto issue_scores = [&](int j) {   // MMA warp, converged
    if (elect_one()) {
      const uint64_t a0 = desc128(base + kOffA + (uint32_t)(j % kStages) * kAStage, 16, 1024);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_f16(tmem, a0 + (uint64_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                 bv_desc + (uint64_t)(((kk >> 2) * (96 * 128) + (kk & 3) * 32) >> 4), kIS, kk > 0);
      umma_commit(bar_s);
    }
    __syncwarp();
  };
  auto issue_grads = [&](int j) {    // MMA warp, converged
    if (elect_one()) {
      // the tile image read MN-major: 16 bank rows per step = two 8-row swizzle atoms (LBO 16384: the second K-block)
      const uint64_t a0 = desc128(base + kOffA + (uint32_t)(j % kStages) * kAStage, 16384, 1024);
      const uint64_t c0 = (j & 1) ? c_desc1 : c_desc0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t acc = (g_started || kk > 0) ? 1u : 0u;
        umma_f16(tmem + 96, a0 + (uint64_t)((kk * 2048) >> 4), c0 + (uint64_t)((kk * 32) >> 4), kIG, acc);
        umma_f16(tmem + 144, a0 + (uint64_t)((64 * 128 + kk * 2048) >> 4), c0 + (uint64_t)((kCImg + kk * 32) >> 4), kIG, acc);
      }
      umma_commit((j & 1) ? bar_g1 : bar_g0);
    }
    __syncwarp();
    g_started = true;
  };

  float ls = 0.f, lt = 0.f;
  uint32_t n_s = 0, n_g0 = 0, n_g1 = 0;   // (workers) waits done on each barrier = its completed phases consumed
  unsigned prev_rec[kRecRegs];            // this thread's records of tile it-1 (their coefficients are cleared at retire)
#pragma unroll
  for (int j = 0; j < kRecRegs; ++j) prev_rec[j] = 0u;

  // one record: its loss terms, and d(loss)/d(score) of its (anchor, row) slot for a negative (dn) and for the positive (dp).
  // Repeats of a slot (sampling with replacement) share the score, so the slot's coefficient is count * dn (+ dp): the
  // sample stage only COUNTS records per slot (native integer atomics) and the product is rounded to bf16 once.
  struct Coef { float dn1, dn2, dp1, dp2; };
  auto eval = [&](unsigned rc, bool with_loss) {
    const unsigned r = rc & 63u, b = (rc >> 6) & 0x3ffu;
    const float s2 = Sd[(size_t)r * kSdPitch + b];               // bank-1 row . v2  (out_v2 direction)
    const float s1 = Sd[(size_t)(64 + r) * kSdPitch + b];        // bank-2 row . v1  (out_v1 direction)
    const float e1 = ex2_approx(s1 * p.k_exp), e2 = ex2_approx(s2 * p.k_exp);
    const float o1 = e1 * p.inv_Z1, o2 = e2 * p.inv_Z2;
    const float rc1 = rcp_approx(o1 + p.c) * p.inv_BT, rc2 = rcp_approx(o2 + p.c) * p.inv_BT;
    if (with_loss) {
      float t1, t2;
      if ((rc >> 31) != 0u) {
        t1 = logf(__fdiv_rn(o1, o1 + p.c));
        t2 = logf(__fdiv_rn(o2, o2 + p.c));
      } else {
        t1 = -log1p_pos(fmaf(o1, p.inv_mPn, p.eps_over_mPn));
        t2 = -log1p_pos(fmaf(o2, p.inv_mPn, p.eps_over_mPn));
      }
      ls += t1;
      lt += t2;
    }
    return Coef{o1 * rc1, o2 * rc2, -p.c * rc1, -p.c * rc2};
  };
  // the slot's coefficients from its final count -> the two operand images (C2^T[b][r]: bank-1 row -> grad_v2[b]; C1^T: bank 2)
  auto store_slot = [&](uint8_t* cimg, unsigned rc, const Coef& k) {
    const unsigned r = rc & 63u, b = (rc >> 6) & 0x3ffu;
    const unsigned n = cnt[b * 64 + r];
    const float nn = (float)(n & (kPosOne - 1u)), np = (float)(n >> 28);
    const uint32_t co = coef_off(b, r);
    *reinterpret_cast<__nv_bfloat16*>(cimg + co) = __float2bfloat16_rn(fmaf(nn, k.dn2, np * k.dp2));
    *reinterpret_cast<__nv_bfloat16*>(cimg + kCImg + co) = __float2bfloat16_rn(fmaf(nn, k.dn1, np * k.dp1));
  };

  // records of the next two tiles ride in registers: a load issued two iterations ahead has landed when it is needed
  // (loops over a thread's record slots stop at the tile's CTA-uniform slot count: one branch skips the unused slots)
  auto slots_of = [&](unsigned cnt_records) { return (int)min((cnt_records + kWorkers - 1) / kWorkers, (unsigned)kRecRegs); };
  auto fetch_records = [&](int j, unsigned* out) {
    if (!worker || j >= ntiles) return;
    const unsigned f0 = offs[j], f1 = offs[j + 1];
    const int kmax = slots_of(f1 - f0);
#pragma unroll
    for (int k = 0; k < kRecRegs; ++k) {
      if (k >= kmax) break;
      const unsigned i = f0 + tid + k * kWorkers;
      if (i < f1) out[k] = __ldg(p.records + i);
    }
  };
  unsigned rec_a[kRecRegs] = {}, rec_b[kRecRegs] = {};
  fetch_records(0, rec_a);
  fetch_records(1, rec_b);
  uint32_t m_g0 = 0, m_g1 = 0;   // (MMA warp) its own phase counts of the gradient barriers
  if (warp == 8 && ntiles > 0) {
    if (elect_one()) {
      load_tile(0);
      if (ntiles > 1) load_tile(1);
      if (ntiles > 2) load_tile(2);
    }
    __syncwarp();
    wait_bar(bar_full, 0);
    tc_fence_after();
    if (active(0)) issue_scores(0);
  }

  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;
  const bool prof = p.prof != 0 && blockIdx.x == 0 && tid == 0;
  auto tick = [&](int k) {
    if (prof) {
      const long long t = clock64();
      pc[k] += t - pt;
      pt = t;
    }
  };
  if (prof) pt = clock64();
  for (int it = 0; it < ntiles; ++it) {
    const bool act = active(it);
    const unsigned n0 = offs[it], n1 = offs[it + 1];
    unsigned rec[kRecRegs];
#pragma unroll
    for (int j = 0; j < kRecRegs; ++j) { rec[j] = rec_a[j]; rec_a[j] = rec_b[j]; }
    fetch_records(it + 2, rec_b);
    if (worker) {
      // ---- (a) scores of tile it: TMEM -> shared memory
      if (act) {
        wait_bar(bar_s, n_s & 1u);
        ++n_s;
        tc_fence_after();
        tick(0);
        const int q = warp & 3, half = warp >> 2;
        const int m = q * 32 + lane, bank = m >> 6, r = m & 63;   // stacked row = TMEM lane; bank-1 rows keep the V2 columns
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(bank * 48 + half * 24);
        uint32_t v[24];
        tmem_ld16(taddr, v);
        tmem_ld8(taddr + 16, v + 16);
        tmem_ld_wait();
        float* dst = Sd + (size_t)(bank * 64 + r) * kSdPitch + half * 24;
#pragma unroll
        for (int i = 0; i < 24; ++i) dst[i] = __uint_as_float(v[i]);
      }
      fence_proxy_async();         // the coefficient stores of tile it-1 (steps b, c of the previous iteration) -> async proxy
      tick(1);
    }
    tc_fence_before();
    bar_sync(1, kThreads);         // #1
    tc_fence_after();
    tick(2);
    Coef ck[kRecRegs];
    unsigned own = 0u;
    const bool dense_tile = n1 - n0 > (unsigned)(kRecRegs * kWorkers);   // more records than the fast path holds (small banks)
    const int jmax = slots_of(n1 - n0);
    uint8_t* cimg_it = sm + kOffC + (uint32_t)(it & 1) * 2 * kCImg;
    if (!worker) {
      // ---- (b, MMA warp) gradient MMAs of tile it-1, score MMAs of tile it+1: both run under the sample stage
      if (act) ++n_s;   // (keeps this warp's phase count of bar_s equal to the workers')
      if (active(it - 1)) issue_grads(it - 1);
      if (it + 1 < ntiles) {
        wait_bar(bar_full + 8 * (uint32_t)((it + 1) % kStages), (uint32_t)(((it + 1) / kStages) & 1));
        tc_fence_after();
        if (active(it + 1)) issue_scores(it + 1);
      }
      // the stage of tile it-1 is free once its gradient MMAs have retired: refill it with tile it+3
      if (active(it - 1)) {
        if ((it - 1) & 1) { wait_bar(bar_g1, m_g1 & 1u); ++m_g1; } else { wait_bar(bar_g0, m_g0 & 1u); ++m_g0; }
      }
      if (it + 3 < ntiles) {
        if (elect_one()) load_tile(it + 3);
        __syncwarp();
      }
    } else if (act && dense_tile) {
      // ---- (b, workers) crowded tile, generic path: count, then (after the barrier) every record stores its slot's value
      for (unsigned i = n0 + tid; i < n1; i += kWorkers) {
        const unsigned rc = __ldg(p.records + i);
        atomicAdd(cnt + ((rc >> 6) & 0x3ffu) * 64 + (rc & 63u), (rc >> 31) ? kPosOne : 1u);
        eval(rc, true);
      }
    } else if (act) {
      // ---- (b, workers) sample stage of tile it, fast path: the first record to count itself on a slot owns the slot
#pragma unroll
      for (int j = 0; j < kRecRegs; ++j) {
        if (j >= jmax) break;
        if (n0 + tid + j * kWorkers < n1) {
          const unsigned rc = rec[j];
          const unsigned old = atomicAdd(cnt + ((rc >> 6) & 0x3ffu) * 64 + (rc & 63u), (rc >> 31) ? kPosOne : 1u);
          ck[j] = eval(rc, true);
          if (old == 0u) own |= 1u << j;
        }
      }
    }
    tick(3);
    if (worker) bar_sync(2, kWorkers);   // #2 (the MMA warp is not part of the count -> store hand-off)
    tick(4);
    if (worker && act) {
      // ---- (c) the counts are final: owners write their slot's coefficients and reset its counter
      if (!dense_tile) {
#pragma unroll
        for (int j = 0; j < kRecRegs; ++j) {
          if (j >= jmax) break;
          if (own & (1u << j)) {
            store_slot(cimg_it, rec[j], ck[j]);
            cnt[((rec[j] >> 6) & 0x3ffu) * 64 + (rec[j] & 63u)] = 0u;
          }
        }
      } else {
        for (unsigned i = n0 + tid; i < n1; i += kWorkers) {   // (repeats store the same value)
          const unsigned rc = __ldg(p.records + i);
          store_slot(cimg_it, rc, eval(rc, false));
        }
      }
    }
    if (dense_tile && worker) {   // the generic path reads the counters and the score dump from every record's thread
      bar_sync(2, kWorkers);
      if (act)
        for (unsigned i = n0 + tid; i < n1; i += kWorkers) {
          const unsigned rc = __ldg(p.records + i);
          cnt[((rc >> 6) & 0x3ffu) * 64 + (rc & 63u)] = 0u;
        }
    }
    tick(5);
    if (worker) {
      // ---- retire tile it-1: its gradient MMAs (issued in step b) are done -> clear its coefficients
      if (active(it - 1)) {
        if ((it - 1) & 1) { wait_bar(bar_g1, n_g1 & 1u); ++n_g1; } else { wait_bar(bar_g0, n_g0 & 1u); ++n_g0; }
        tick(6);
        uint8_t* cimg = sm + kOffC + (uint32_t)((it - 1) & 1) * 2 * kCImg;
        const unsigned z0 = offs[it - 1], z1 = offs[it];
        auto clear = [&](unsigned rc) {
          const uint32_t co = coef_off((rc >> 6) & 0x3ffu, rc & 63u);
          *reinterpret_cast<unsigned short*>(cimg + co) = 0;
          *reinterpret_cast<unsigned short*>(cimg + kCImg + co) = 0;
        };
const int zmax = slots_of(z1 - z0);
#pragma unroll
        for (int j = 0; j < kRecRegs; ++j) {
          if (j >= zmax) break;
          if (z0 + tid + j * kWorkers < z1) clear(prev_rec[j]);
        }
        for (unsigned i = z0 + tid + kRecRegs * kWorkers; i < z1; i += kWorkers) clear(__ldg(p.records + i));
      }
#pragma unroll
      for (int j = 0; j < kRecRegs; ++j) prev_rec[j] = rec[j];
    }
    tick(7);
  }
  if (prof)
    printf("crd_tc_stream CTA 0: %d tiles; cycles wait_scores %lld dump+fence %lld bar1 %lld sample %lld bar2 %lld store %lld wait_grads %lld clear %lld\n",
           ntiles, pc[0], pc[1], pc[2], pc[3], pc[4], pc[5], pc[6], pc[7]);
  // ---- the last tile's gradient MMAs, then everything issued has to retire before the accumulators are read
  if (worker) fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) {
    if (active(ntiles - 1)) issue_grads(ntiles - 1);
    if (elect_one()) umma_commit(bar_s);
    __syncwarp();
  }
  wait_bar(bar_s, n_s & 1u);
  tc_fence_after();
  const bool any_grad = offs[ntiles] != offs[0];
  // ---- flush: G2^T / G1^T (lane = feature e) -> partial[cta][b][0..127 grad_v1 | 128..255 grad_v2]
  if (worker) {
    const int q = warp & 3, which = warp >> 2;            // which 0: G2^T (cols 96..143), 1: G1^T (cols 144..191)
    const int e = q * 32 + lane;
    float* part = p.partial + (size_t)blockIdx.x * p.B * 256 + (which == 0 ? 128 : 0) + e;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 96u + 48u * which;
#pragma unroll
    for (int c0 = 0; c0 < 48; c0 += 16) {
      uint32_t v[16];
      if (any_grad) { tmem_ld16(taddr + c0, v); tmem_ld_wait(); }
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.B) part[(size_t)(c0 + i) * 256] = any_grad ? __uint_as_float(v[i]) : 0.f;
    }
  }
  for (int o = 16; o >= 1; o >>= 1) {
    ls += __shfl_xor_sync(0xffffffffu, ls, o);
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
  }
  if (lane == 0 && worker) {
    float* lp = p.loss_part + (size_t)blockIdx.x * 32;
    lp[2 * warp] = ls; lp[2 * warp + 1] = lt;
    lp[2 * (warp + 8)] = 0.f; lp[2 * (warp + 8) + 1] = 0.f;   // ts_finalize sums 16 slots per CTA
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256u);
  }
}

}  // namespace tc
