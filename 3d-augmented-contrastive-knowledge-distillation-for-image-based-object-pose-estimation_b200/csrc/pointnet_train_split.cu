// pointnet_train_split.cu -- fp32-ACCURATE train-mode forward of the PointNet encoder on tcgen05 tensor cores
// (ShapeEncoderPC.forward under model.train(), auxiliary/model.py:174-180, differentiated at training.py:75).
//
// Why it exists: the reference trains in fp32.  Train-mode gradients are ROUTED by discrete decisions -- the arg-max
// point of every (cloud, channel) and the ReLU gates -- so a bf16 forward (4e-3 relative noise on the conv outputs)
// flips about 1% of the arg-max decisions and the parameter gradients then differ from the reference's by 12-18%
// although every feature is within 1e-2.  Here every tensor-core product is evaluated with both operands split into
// two fp16 halves, x = hi + lo (22 mantissa bits), as three MMAs with fp32 accumulation in TMEM:
//     W.h  ~=  Whi.hhi + Wlo.hhi + Whi.hlo          (the dropped lo.lo term is 2^-22 relative)
// which brings the conv outputs to ~2e-7 of the fp32 reference (measured against fp64 on the CPU,
// profiles/experiments/r2_pn_split_recipe.py), i.e. the routing is the reference's up to fp32 rounding.  The weights
// are pre-scaled by 2^8 (exact) so that the lo halves stay in fp16's normal range; the scale is taken out again in
// the epilogues.  The layer-3 epilogue keeps the EXACT fp32 maximum and its first arg-max point (no mantissa bits
// are borrowed for the index as in the bf16 recipe).
//
// Work decomposition: a unit = 128 consecutive points of one cloud (= one h2 tile of the train context), one
// persistent CTA per SM.  Shared memory (216 KB): ring of five 16 KB W3 pieces (128 channels x 64 k; a slab is four
// pieces: hi k0, hi k1, lo k0, lo k1), h2 hi/lo (64 KB), h1 hi/lo (32 KB), W2 hi/lo (32 KB).  TMEM: four 128-column
// accumulator slots shared by all jobs (per unit one layer-2 job and F/128 layer-3 jobs of 24 MMAs each).
// Warp roles (512 threads): w0 bulk-copy producer | w1 MMA issuer | w2 TMEM allocator | w3 idle |
//                           w4-7 layer-3 epilogue (max / arg-max / BN3 sums) | w8-15 front end (layer 1 on CUDA cores in
//                           fp32, layer-2 epilogue: BN2, ReLU, split, h2 kept in bf16 for the backward).
#include <cuda_fp16.h>

#include "pointnet_common.cuh"
#include "pointnet_train.cuh"

namespace crdpn {
namespace pn {
namespace sp {

constexpr int kUnit = 128;
constexpr uint32_t kPiece = 16384;  // 128 rows x 64 k of a 16-bit type, K-major SWIZZLE_128B
constexpr int kRing = 5;
constexpr uint32_t kOffRing = 0;
constexpr uint32_t kOffH2 = kOffRing + kRing * kPiece;  // hi k0 | hi k1 | lo k0 | lo k1
constexpr uint32_t kOffH1 = kOffH2 + 4 * kPiece;        // hi | lo
constexpr uint32_t kOffW2 = kOffH1 + 2 * kPiece;        // hi | lo
constexpr uint32_t kOffPar = kOffW2 + 2 * kPiece;       // W1p[64][4] f32, sh2[128], sc2[128] * 2^-8
constexpr uint32_t kOffBar = kOffPar + 2048;
constexpr int kNumBars = 32;
constexpr uint32_t kSmemBytes = kOffBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemAlloc = kSmemBytes + 1024;
constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;

enum Bar : int {
  W3_FULL = 0,     // [5] bulk copy landed
  W3_EMPTY = 5,    // [5] commit: the MMAs reading the piece have retired
  H1_FULL = 10,    // 256 arrivals
  H1_EMPTY = 11,   // commit
  H2_FULL = 12,    // 256 arrivals
  H2_EMPTY = 13,   // commit after the last slab of a unit
  A2_FULL = 14,    // commit: layer-2 job of unit u landed (phase u & 1)
  S_FULL = 15,     // [4] commit: a layer-3 job landed in slot i
  ACC_EMPTY = 19   // [4] 256 arrivals: slot i drained (every job, in order; only the MMA issuer waits on it)
};

// kind::f16 with F16 (not BF16) operands, fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
// x (>= 0 after the ReLU, or any sign) -> fp16 hi / lo halves of two values, low half <- a
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(a, 65504.f);
  b = fminf(b, 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 back = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Params {
  const float* x;  // [B,3,P]
  int B, P, F;
  int tiles_per_cloud;  // 128-point tiles per cloud (TrainCtx::tiles2)
  int total_units;
  const char* packed;       // W2 hi | W2 lo | per slab: hi k0 | hi k1 | lo k0 | lo k1   (fp16, x 2^8)
  const float* train_par;   // W1p[64][4], sh2[128], sc2[128]
  char* h2img;              // bf16 tiles of h2 for the backward
  unsigned long long* enc64;
  double *sum3, *sq3;
};

template <int NSLAB>
__global__ void __launch_bounds__(kThreads, 1) pointnet_fwd_train_split_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + kOffBar;
  auto bar = [&](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kOffBar + kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int u_begin = (int)(((long long)p.total_units * blockIdx.x) / G);
  const int u_end = (int)(((long long)p.total_units * (blockIdx.x + 1)) / G);
  const int NU = u_end - u_begin;
  // job numbering (identical in every role): L2(0) = 0; unit u's slabs follow at u(NS+1)+1+s, except that the layer-2
  // job of unit u+1 is slotted in right before slab kL2At of unit u
  constexpr int kL2At = NSLAB >= 2 ? NSLAB - 2 : 0;
  auto job_l2 = [&](int u) -> uint32_t { return u == 0 ? 0u : (uint32_t)((u - 1) * (NSLAB + 1) + 1 + kL2At); };
  auto job_s = [&](int u, int s) -> uint32_t {
    return (uint32_t)(u * (NSLAB + 1) + 1 + s + ((s >= kL2At && u + 1 < NU) ? 1 : 0));
  };

  {
    float* spar = reinterpret_cast<float*>(sm + kOffPar);
    for (int i = threadIdx.x; i < 512; i += kThreads) spar[i] = i >= 384 ? p.train_par[i] * kWInv : p.train_par[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(bar(W3_FULL + i), 1); mbar_init(bar(W3_EMPTY + i), 1); }
    mbar_init(bar(H1_FULL), 256); mbar_init(bar(H1_EMPTY), 1);
    mbar_init(bar(H2_FULL), 256); mbar_init(bar(H2_EMPTY), 1);
    mbar_init(bar(A2_FULL), 1);
    for (int i = 0; i < 4; ++i) { mbar_init(bar(S_FULL + i), 1); mbar_init(bar(ACC_EMPTY + i), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =========================== bulk-copy producer ===========================
    if (lane == 0 && NU > 0) {
      // W2 hi | lo: its arrival is observed through the first ring barrier's transaction count
      const char* w3 = p.packed + 2 * kPiece;
      uint32_t n = 0;
      for (int u = 0; u < NU; ++u) {
        for (int s = 0; s < NSLAB; ++s) {
#pragma unroll 1
          for (int pc = 0; pc < 4; ++pc, ++n) {
            const uint32_t stage = n % kRing, use = n / kRing;
            mbar_wait(bar(W3_EMPTY + stage), (use & 1u) ^ 1u);
            const bool first = n == 0;
            mbar_expect_tx(bar(W3_FULL + stage), kPiece + (first ? 2 * kPiece : 0u));
            if (first) {
              bulk_g2s(base + kOffW2, p.packed, kPiece, bar(W3_FULL + stage));
              bulk_g2s(base + kOffW2 + kPiece, p.packed + kPiece, kPiece, bar(W3_FULL + stage));
            }
            const uint32_t dst = base + kOffRing + stage * kPiece;
            const char* src = w3 + ((size_t)s * 4 + pc) * kPiece;
            bulk_g2s(dst, src, 8192u, bar(W3_FULL + stage));
            bulk_g2s(dst + 8192u, src + 8192, 8192u, bar(W3_FULL + stage));
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (converged warp, one elected lane issues) ===========================
    if (NU > 0) {
      constexpr uint32_t kI = idesc_f16(128, 128);
      mbar_wait(bar(W3_FULL + 0), 0);  // W2 rides on the first ring transaction (phase 0 of stage 0)
      uint32_t j = 0, piece_n = 0;
      auto issue_layer2 = [&](int u) {  // D2[point][channel] = h1 . W2^T as hi.hi + lo.hi + hi.lo, K = 64
        const uint32_t uph = (uint32_t)u & 1u, slot = j & 3u;
        mbar_wait(bar(H1_FULL), uph);
        mbar_wait(bar(ACC_EMPTY + slot), ((j >> 2) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_hi = umma_desc_sw128(base + kOffH1), a_lo = umma_desc_sw128(base + kOffH1 + kPiece);
          const uint64_t b_hi = umma_desc_sw128(base + kOffW2), b_lo = umma_desc_sw128(base + kOffW2 + kPiece);
          const uint32_t d = tmem + 128u * slot;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2u * k, b_hi + 2u * k, kI, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_lo + 2u * k, b_hi + 2u * k, kI, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2u * k, b_lo + 2u * k, kI, 1u);
          umma_commit(bar(A2_FULL));
          umma_commit(bar(H1_EMPTY));
        }
        __syncwarp();
        ++j;
      };
      issue_layer2(0);
      for (int u = 0; u < NU; ++u) {
        const uint32_t uph = (uint32_t)u & 1u;
        for (int s = 0; s < NSLAB; ++s) {
          if (s == kL2At && u + 1 < NU) issue_layer2(u + 1);
          const uint32_t slot = j & 3u;
          if (s == 0) mbar_wait(bar(H2_FULL), uph);
          mbar_wait(bar(ACC_EMPTY + slot), ((j >> 2) & 1u) ^ 1u);
          const uint32_t d = tmem + 128u * slot;
#pragma unroll 1
          for (int pc = 0; pc < 4; ++pc, ++piece_n) {  // hi k0 | hi k1 | lo k0 | lo k1
            const uint32_t stage = piece_n % kRing;
            // stage 0 / phase 0 was already observed above (re-waiting on a completed phase returns at once)
            mbar_wait(bar(W3_FULL + stage), (piece_n / kRing) & 1u);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t kb = (uint32_t)pc & 1u;
              const uint64_t a = umma_desc_sw128(base + kOffRing + stage * kPiece);
              const uint64_t b_hi = umma_desc_sw128(base + kOffH2 + kb * kPiece);
              const uint64_t b_lo = umma_desc_sw128(base + kOffH2 + (2u + kb) * kPiece);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d, a + 2u * k, b_hi + 2u * k, kI, (pc > 0 || k > 0) ? 1u : 0u);
              if (pc < 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, a + 2u * k, b_lo + 2u * k, kI, 1u);
              }
              umma_commit(bar(W3_EMPTY + stage));
              if (pc == 3) {
                umma_commit(bar(S_FULL + slot));
                if (s == NSLAB - 1) umma_commit(bar(H2_EMPTY));
              }
            }
            __syncwarp();
          }
          ++j;
        }
      }
      // the epilogue has drained the last job => every MMA and every commit queued behind it has retired
      mbar_wait(bar(ACC_EMPTY + ((j - 1u) & 3u)), ((j - 1u) >> 2) & 1u);
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================== layer-3 epilogue: exact max, first arg-max point, BN3 sums =====================
    const int q = warp & 3;
    float rmax[NSLAB];
    int ridx[NSLAB];
    double rs[NSLAB], rq[NSLAB];
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) { rmax[s] = -INFINITY; ridx[s] = 0; rs[s] = 0.0; rq[s] = 0.0; }
    int cur_cloud = -1;
    uint32_t seen = 0;  // bit i: parity of the layer-3 jobs already drained from slot i
    auto flush = [&](int cloud) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const unsigned long long key = ((unsigned long long)enc_ordered(rmax[s] * kWInv) << 32) |
                                       (unsigned long long)(0xffffffffu - (uint32_t)ridx[s]);
        atomicMax(p.enc64 + (size_t)cloud * p.F + s * 128 + q * 32 + lane, key);
        rmax[s] = -INFINITY;
        ridx[s] = 0;
      }
    };
    for (int u = 0; u < NU; ++u) {
      const int unit = u_begin + u;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnit;
      const int ndup = p_base + kUnit > p.P ? (p_base + kUnit - p.P < kUnit ? p_base + kUnit - p.P : kUnit) : 0;
      if (cloud != cur_cloud) {
        if (cur_cloud >= 0) flush(cur_cloud);
        cur_cloud = cloud;
      }
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const uint32_t j = job_s(u, s), slot = j & 3u;
        mbar_wait(bar(S_FULL + slot), (seen >> slot) & 1u);
        seen ^= 1u << slot;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 128u * slot;
        float m = rmax[s];
        int mi = ridx[s];
        unsigned long long sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;
        uint32_t ra[16], rb[16];
        auto chunk = [&](const uint32_t* cur, int c) {
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = __uint_as_float(cur[i]);
          const float t0 = max3(w[0], w[1], w[2]), t1 = max3(w[3], w[4], w[5]), t2 = max3(w[6], w[7], w[8]);
          const float t3 = max3(w[9], w[10], w[11]), t4 = max3(w[12], w[13], w[14]);
          const float cm = fmaxf(max3(t0, t1, t2), max3(t3, t4, w[15]));
          if (__any_sync(0xffffffffu, cm > m)) {  // rare once a few chunks of the cloud have been seen
            if (cm > m) {
              int pos = 15;
#pragma unroll
              for (int i = 14; i >= 0; --i) pos = (w[i] == cm) ? i : pos;
              m = cm;
              mi = p_base + 16 * c + pos;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            add2(sa, cur[i], cur[i + 1]);     add2(sb, cur[i + 2], cur[i + 3]);
            sq2(qa, cur[i], cur[i + 1]);      sq2(qb, cur[i + 2], cur[i + 3]);
          }
        };
        tmem_ld16(taddr, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          tmem_ld16(taddr + 16u * (uint32_t)(c + 1), rb);
          chunk(ra, c);
          tmem_ld_wait();
          if (c + 2 < 8) tmem_ld16(taddr + 16u * (uint32_t)(c + 2), ra);
          chunk(rb, c + 1);
          tmem_ld_wait();
        }
        const float ylast = __uint_as_float(rb[15]);  // column 127: the last real point whenever ndup > 0
        tc_fence_before();
        mbar_arrive_n(bar(ACC_EMPTY + slot), 2u);  // 128 epilogue threads stand in for the slot's 256 arrivals
        if (ndup < kUnit) {  // a tile made of padding only (odd number of 128-point tiles) adds nothing
          rs[s] += (double)(pair_sum(sa) + pair_sum(sb)) - (double)ndup * (double)ylast;
          rq[s] += (double)(pair_sum(qa) + pair_sum(qb)) - (double)ndup * (double)ylast * (double)ylast;
        }
        rmax[s] = m;
        ridx[s] = mi < p.P ? mi : p.P - 1;
      }
    }
    if (cur_cloud >= 0) flush(cur_cloud);
    if (NU > 0) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        atomicAdd(p.sum3 + s * 128 + q * 32 + lane, rs[s] * (double)kWInv);
        atomicAdd(p.sq3 + s * 128 + q * 32 + lane, rq[s] * (double)(kWInv * kWInv));
      }
    }
  } else if (warp >= 8) {
    // =========================== front end: layer 1 (fp32) + layer-2 epilogue ====================================
    // thread (t, g): point row t of the unit (also its TMEM lane), channel half g (layer 1: 32 of 64, layer 2: 64 of 128)
    const int g = warp >= 12 ? 1 : 0;
    const int t = threadIdx.x & 127;
    const int q = warp & 3;
    const float4* w1p = reinterpret_cast<const float4*>(sm + kOffPar);
    const float* sh2 = reinterpret_cast<const float*>(sm + kOffPar + 1024);
    const float* sc2 = reinterpret_cast<const float*>(sm + kOffPar + 1536);

    auto layer1 = [&](int u) {
      const int unit = u_begin + u;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnit;
      const float* xc = p.x + (size_t)cloud * 3 * p.P;
      int pt = p_base + t;
      pt = pt < p.P ? pt : p.P - 1;  // ragged tail: repeat the last real point
      const float x0 = __ldg(xc + pt), x1 = __ldg(xc + p.P + pt), x2 = __ldg(xc + 2 * p.P + pt);
      mbar_wait(bar(H1_EMPTY), ((uint32_t)u & 1u) ^ 1u);
      uint8_t* dhi = sm + kOffH1;
      uint8_t* dlo = sm + kOffH1 + kPiece;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        const int ch = g * 32 + cg * 8;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 wa = w1p[ch + 2 * jj], wb = w1p[ch + 2 * jj + 1];
          const float va = fmaxf(fmaf(wa.x, x0, fmaf(wa.y, x1, fmaf(wa.z, x2, wa.w))), 0.f);
          const float vb = fmaxf(fmaf(wb.x, x0, fmaf(wb.y, x1, fmaf(wb.z, x2, wb.w))), 0.f);
          split2(va, vb, hi[jj], lo[jj]);
        }
        const uint32_t off = sw128_off(t, ch);
        *reinterpret_cast<uint4*>(dhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      mbar_arrive(bar(H1_FULL));
    };

    if (NU > 0) layer1(0);
    for (int u = 0; u < NU; ++u) {
      const uint32_t uph = (uint32_t)u & 1u;
      const uint32_t j = job_l2(u), slot = j & 3u;
      mbar_wait(bar(A2_FULL), uph);
      tc_fence_after();
      uint32_t hi[32], lo[32];  // this point's 64 layer-2 channels: BN2 + ReLU, split into fp16 halves
      char* gt = p.h2img + (size_t)(u_begin + u) * kTileBytes + (size_t)g * kKBlockBytes;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 128u * slot + 64u * (uint32_t)g;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + 32u * c, r);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int ch = g * 64 + c * 32 + g8 * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(__uint_as_float(r[g8 * 8 + e]), sc2[ch + e], sh2[ch + e]), 0.f);
          uint4 o;  // bf16 copy for the backward (gates and dense terms), same swizzled tile image as before
          o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(gt + sw128_off(t, c * 32 + g8 * 8)) = o;
#pragma unroll
          for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hi[c * 16 + g8 * 4 + e], lo[c * 16 + g8 * 4 + e]);
        }
      }
      tc_fence_before();
      mbar_arrive(bar(ACC_EMPTY + slot));  // the values live in registers now
      mbar_wait(bar(H2_EMPTY), uph ^ 1u);  // layer 3 of the previous unit has finished reading h2
      uint8_t* dhi = sm + kOffH2 + (uint32_t)g * kPiece;
      uint8_t* dlo = sm + kOffH2 + (2u + (uint32_t)g) * kPiece;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = sw128_off(t, i * 8);
        *reinterpret_cast<uint4*>(dhi + off) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
        *reinterpret_cast<uint4*>(dlo + off) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
      }
      fence_proxy_async();
      mbar_arrive(bar(H2_FULL));
      if (u + 1 < NU) layer1(u + 1);  // overlaps with layer 3 of unit u on the tensor pipe
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp16 hi / lo operand images (x 2^8): W2 [128 x 64] and sign(gamma3) * W3 [F x 128] in pieces
__global__ void __launch_bounds__(256) pn_pack_train_split_kernel(const float* __restrict__ c2w, const float* __restrict__ c3w,
                                                                  const float* __restrict__ g3, int F, char* __restrict__ packed) {
  const int nW2 = 128 * 64, nW3 = F * 128;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nW2 + nW3; i += gridDim.x * blockDim.x) {
    float v;
    char *hi_at, *lo_at;
    if (i < nW2) {
      const int c = i / 64, k = i % 64;
      v = c2w[i] * kWScale;
      hi_at = packed + sw128_off(c, k);
      lo_at = hi_at + kPiece;
    } else {
      const int jx = i - nW2;
      const int c = jx / 128, k = jx % 128;
      v = c3w[jx] * (g3[c] >= 0.f ? kWScale : -kWScale);
      char* slab = packed + 2 * kPiece + (size_t)(c / 128) * (4 * kPiece);
      hi_at = slab + (size_t)(k / 64) * kPiece + sw128_off(c % 128, k % 64);
      lo_at = hi_at + 2 * kPiece;
    }
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    *reinterpret_cast<__half*>(hi_at) = h;
    *reinterpret_cast<__half*>(lo_at) = __float2half_rn(v - __half2float(h));
  }
}

// ---------------------------------------------------------------------------------------------------------
// BN2 statistics pass in the same split arithmetic: per 256-point unit, h1 = relu(bn1(conv1 x)) in fp32 -> fp16 hi / lo
// operand images, D[channel][point] = W2 . h1^T as three MMA groups (M = 128 channels, N = 256 points, K = 64), each
// thread sums its channel's 128 columns.  Two h1 buffers / TMEM slots: the MMAs of unit u overlap the sums of unit u-1.
struct Stats2Params {
  const float* x;
  int B, P;
  int tiles_per_cloud, total_units;  // 256-point units
  const char* packed;      // W2 hi | W2 lo
  const float* train_par;  // W1p[64][4]
  double *sum2, *sq2;
};
constexpr uint32_t kS2OffW2 = 0, kS2OffH1 = 2 * kPiece, kS2OffPar = kS2OffH1 + 2 * 65536, kS2OffBar = kS2OffPar + 1024;
constexpr uint32_t kS2Smem = kS2OffBar + 64 + 1024;

__global__ void __launch_bounds__(256, 1) pn_stats2_split_kernel(const Stats2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + kS2OffBar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kS2OffBar + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int u_begin = (int)(((long long)p.total_units * blockIdx.x) / G);
  const int u_end = (int)(((long long)p.total_units * (blockIdx.x + 1)) / G);
  const int NU = u_end - u_begin;

  for (int i = tid; i < 2048; i += 256)
    reinterpret_cast<uint4*>(sm + kS2OffW2)[i] = reinterpret_cast<const uint4*>(p.packed)[i];
  reinterpret_cast<float*>(sm + kS2OffPar)[tid] = p.train_par[tid];
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float4* w1p = reinterpret_cast<const float4*>(sm + kS2OffPar);
  const int q = warp & 3, hf = warp >> 2;
  constexpr uint32_t kI = idesc_f16(128, 256);
  double rs = 0.0, rq = 0.0;

  for (int it = 0; it <= NU; ++it) {
    if (it < NU) {
      const int unit = u_begin + it;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnitPts;
      const float* xc = p.x + (size_t)cloud * 3 * p.P;
      int pt = p_base + tid;
      pt = pt < p.P ? pt : p.P - 1;
      const float x0 = __ldg(xc + pt), x1 = __ldg(xc + p.P + pt), x2 = __ldg(xc + 2 * p.P + pt);
      uint8_t* dhi = sm + kS2OffH1 + (it & 1) * 65536;
      uint8_t* dlo = dhi + 32768;
#pragma unroll
      for (int cg = 0; cg < 8; ++cg) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 wa = w1p[cg * 8 + 2 * jj], wb = w1p[cg * 8 + 2 * jj + 1];
          const float va = fmaxf(fmaf(wa.x, x0, fmaf(wa.y, x1, fmaf(wa.z, x2, wa.w))), 0.f);
          const float vb = fmaxf(fmaf(wb.x, x0, fmaf(wb.y, x1, fmaf(wb.z, x2, wb.w))), 0.f);
          split2(va, vb, hi[jj], lo[jj]);
        }
        const uint32_t off = sw128_off(tid, cg * 8);
        *reinterpret_cast<uint4*>(dhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (it < NU && warp == 0) {
      if (elect_one()) {
        const uint64_t a_hi = umma_desc_sw128(base + kS2OffW2), a_lo = umma_desc_sw128(base + kS2OffW2 + kPiece);
        const uint64_t b_hi = umma_desc_sw128(base + kS2OffH1 + (it & 1) * 65536);
        const uint64_t b_lo = umma_desc_sw128(base + kS2OffH1 + (it & 1) * 65536 + 32768);
        const uint32_t d = tmem + 256u * (it & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2u * k, b_hi + 2u * k, kI, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, a_lo + 2u * k, b_hi + 2u * k, kI, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, a_hi + 2u * k, b_lo + 2u * k, kI, 1u);
        umma_commit(bar0 + 8u * (it & 1));
      }
      __syncwarp();
    }
    if (it > 0) {
      const int pu = it - 1;
      const uint32_t slot = (uint32_t)pu & 1u;
      mbar_wait(bar0 + 8u * slot, ((uint32_t)pu >> 1) & 1u);
      tc_fence_after();
      const int unit = u_begin + pu;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnitPts;
      const int ndup = p_base + kUnitPts > p.P ? p_base + kUnitPts - p.P : 0;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + 256u * slot;
      unsigned long long s2 = 0ull, q2 = 0ull;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(trow + 128u * hf + 32u * c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) { add2(s2, r[i], r[i + 1]); sq2(q2, r[i], r[i + 1]); }
      }
      double fs = (double)pair_sum(s2), fq = (double)pair_sum(q2);
      if (ndup > 0) {  // padded columns repeat the last real point: take them out again
        const int valid = kUnitPts - ndup, lo = 128 * hf;
        const int dups = lo + 128 - (valid > lo ? valid : lo);
        if (dups > 0) {
          uint32_t yl;
          tmem_ld1(trow + 255u, yl);
          tmem_ld_wait();
          const double y = (double)__uint_as_float(yl);
          fs -= (double)dups * y;
          fq -= (double)dups * y * y;
        }
      }
      rs += fs;
      rq += fq;
    }
  }
  if (NU > 0) {
    atomicAdd(p.sum2 + q * 32 + lane, rs * (double)kWInv);
    atomicAdd(p.sq2 + q * 32 + lane, rq * (double)(kWInv * kWInv));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

template <int NSLAB>
static int launch_split(const Params& fp, int grid, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  if (!attr_set[device]) {
    CRDPN_CUDA(cudaFuncSetAttribute(pointnet_fwd_train_split_kernel<NSLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kSmemAlloc));
    attr_set[device] = true;
  }
  {
    ScopedKernelTimer tm(CRDPN_K_POINTNET_FWD, st);
    pointnet_fwd_train_split_kernel<NSLAB><<<grid, kThreads, kSmemAlloc, st>>>(fp);
  }
  CRDPN_LAUNCH_CHECK("pointnet_fwd_train_split_kernel");
  return CRDPN_OK;
}

}  // namespace sp

// phases 0-2 of the train forward in split (fp32-accurate) arithmetic; called by crdpn_pointnet_forward_train_phased
int split_pack(const float* conv2_w, const float* conv3_w, const float* bn3_w, int F, char* packed, int sms, cudaStream_t st) {
  sp::pn_pack_train_split_kernel<<<sms, 256, 0, st>>>(conv2_w, conv3_w, bn3_w, F, packed);
  CRDPN_LAUNCH_CHECK("pn_pack_train_split_kernel");
  return CRDPN_OK;
}

int split_stats2(const float* x, int B, int P, const char* packed, const float* train_par, double* sum2, double* sq2, int sms,
                 cudaStream_t st) {
  static bool attr_set[64] = {false};
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  if (!attr_set[device]) {
    CRDPN_CUDA(cudaFuncSetAttribute(sp::pn_stats2_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp::kS2Smem));
    attr_set[device] = true;
  }
  const int tiles_per_cloud = (P + kUnitPts - 1) / kUnitPts;
  const int total_units = B * tiles_per_cloud;
  const int grid = total_units < sms ? total_units : sms;
  sp::Stats2Params a{x, B, P, tiles_per_cloud, total_units, packed, train_par, sum2, sq2};
  sp::pn_stats2_split_kernel<<<grid, 256, sp::kS2Smem, st>>>(a);
  CRDPN_LAUNCH_CHECK("pn_stats2_split_kernel");
  return CRDPN_OK;
}

int split_forward(const float* x, int B, int P, int F, const char* packed, const float* train_par, char* h2img,
                  unsigned long long* enc64, double* sum3, double* sq3, int sms, cudaStream_t st) {
  sp::Params a;
  a.x = x; a.B = B; a.P = P; a.F = F;
  a.tiles_per_cloud = 2 * ((P + 255) / 256);
  a.total_units = B * a.tiles_per_cloud;
  a.packed = packed; a.train_par = train_par; a.h2img = h2img; a.enc64 = enc64; a.sum3 = sum3; a.sq3 = sq3;
  const int grid = a.total_units < sms ? a.total_units : sms;
  switch (F / 128) {
    case 1: return sp::launch_split<1>(a, grid, st);
    case 2: return sp::launch_split<2>(a, grid, st);
    case 4: return sp::launch_split<4>(a, grid, st);
    default: return sp::launch_split<8>(a, grid, st);
  }
}

}  // namespace pn
}  // namespace crdpn
