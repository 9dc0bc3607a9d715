// pointnet_kernels.cu -- the teacher's PointNet encoder (ShapeEncoderPC.forward, auxiliary/model.py:174-180)
// as ONE fused sm_100a kernel: shared MLP 3->64->128->F on points + channel-wise max over points.
//
//   layer 1 (3->64,  0.07% of FLOPs)  CUDA cores, fp32 in, BN1 folded, ReLU, bf16 out -> smem (UMMA operand)
//   layer 2 (64->128, 6%)             tcgen05.mma, M = 128 points, N = 128 channels, K = 64; epilogue adds the
//                                     folded bias, ReLU, bf16 -> smem as the K-major B operand of layer 3
//   layer 3 (128->F, 94%)             tcgen05.mma "swap-AB": M = 128 channels (one W3 slab), N = 256 points,
//                                     K = 128; accumulator lanes are channels, so the max over points is a
//                                     per-thread reduction over TMEM columns; BN3 (scale folded into W3, shift
//                                     added after the max) commutes with the max because max(a*y) with the sign
//                                     of a folded into W3 is just max of the folded product.
// Nothing of size B*F*P is ever written: the [B,1024,2500] activation the stock path materialises (1.64 GB)
// lives only in TMEM, 128x128 fp32 at a time.
//
// Work decomposition: a unit = 256 consecutive points of one cloud (two 128-point halves; the ragged tail is
// padded by repeating the last real point -- max is idempotent).  Units are cut into equal contiguous ranges,
// one per SM (persistent CTAs).  Each W3 slab (128 channels x 128 k, 32 KB bf16, pre-swizzled image in global
// memory, L2 resident) is streamed by 1-D bulk async copies through a 3-stage ring and used for both halves.
//
// Warp roles, TMEM plan and the job pipeline: see pointnet_fwd_kernel_v2 below (the first version, N = 128 MMAs with
// one epilogue group per half, was limited by the 128 B/clk shared-memory port and has been removed: git history).

#include "pointnet_common.cuh"

namespace crdpn {
namespace pn {


// ---------------------------------------------------------------------------------------------------------
// v2: layer 3 issues N = 256 MMAs (both halves of a unit in one instruction: 12 KB of shared-memory operands per
// 128-cycle instruction = 96 B/clk instead of the 128 B/clk that N = 128 needs -- the shared-memory port was the
// limiter of v1).  TMEM = ring of two 256-column slots shared by ALL accumulator jobs: per unit one layer-2 job
// (half 0 -> columns [0,128), half 1 -> [128,256) of its slot, drained by the two front-end groups) and NSLAB
// layer-3 jobs (drained by the epilogue warps).  The layer-2 job of unit u+1 is issued before the last slab of unit
// u; the front end pulls it into registers (bias, ReLU, bf16) right away and only the 16 STS.128 per thread wait for
// h2 to be released.
// Warp roles (512 threads): w0 producer | w1 MMA issuer | w2 TMEM allocator | w3 idle | w4-7 layer-3 epilogue |
//                           w8-11 front end for half 0 | w12-15 front end for half 1
// ---------------------------------------------------------------------------------------------------------
namespace v2 {
enum Bar2 : int {
  W3_FULL = 0,    // [3]
  W3_EMPTY = 3,   // [3]
  H1_FULL = 6,    // [2] 128 arrivals
  H1_EMPTY = 8,   // [2] commit
  H2_FULL = 10,   // [2] 128 arrivals
  H2_EMPTY = 12,  // [1] commit
  S_FULL = 13,    // [2] commit: a layer-3 job landed in slot i (phase = number of layer-3 jobs seen on that slot)
  ACC_EMPTY = 15, // [2] 256 arrivals: slot i drained (every job, in order -- only the MMA issuer waits on it)
  W2_FULL = 17,   // [1]
  A2_FULL = 18    // [1] commit: the layer-2 job of unit u landed (phase = u)
  // NOTE: a waiter may only use parity waits on a barrier whose EVERY phase it observes; the front end skips the
  // layer-3 jobs and the epilogue skips the layer-2 jobs, hence the separate FULL barriers per job type.
};
constexpr uint32_t kIdescN128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescN256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kH2KBlockBytes = 32768;  // 256 rows x 64 k bf16

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
}  // namespace v2

template <int NSLAB, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) pointnet_fwd_kernel_v2(const FwdParams p) {
  using namespace v2;
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
  // job numbering (identical in every role): L2(0)=0; unit u's slabs follow at u(NS+1)+1+s, except that the
  // layer-2 job of unit u+1 is slotted in right before slab kL2At of unit u (two slabs before the end, so its
  // epilogue has two slab-times to pull the accumulator into registers before h2 is released)
  constexpr int kL2At = NSLAB >= 2 ? NSLAB - 2 : 0;
  auto job_l2 = [&](int u) -> uint32_t { return u == 0 ? 0u : (uint32_t)((u - 1) * (NSLAB + 1) + 1 + kL2At); };
  auto job_s = [&](int u, int s) -> uint32_t {
    return (uint32_t)(u * (NSLAB + 1) + 1 + s + ((s >= kL2At && u + 1 < NU) ? 1 : 0));
  };

  {
    const float* par = TRAIN ? p.train_par : reinterpret_cast<const float*>(p.packed + packed_off_par(p.F));
    float* spar = reinterpret_cast<float*>(sm + kOffPar);
    for (int i = threadIdx.x; i < (int)(kParBytes / 4); i += kThreads) spar[i] = par[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(W3_FULL + i), 1); mbar_init(bar(W3_EMPTY + i), 1); }
    for (int h = 0; h < 2; ++h) {
      mbar_init(bar(H1_FULL + h), kHalfPts); mbar_init(bar(H1_EMPTY + h), 1);
      mbar_init(bar(H2_FULL + h), kHalfPts);
      mbar_init(bar(S_FULL + h), 1);         mbar_init(bar(ACC_EMPTY + h), 2 * kHalfPts);
    }
    mbar_init(bar(H2_EMPTY), 1);
    mbar_init(bar(W2_FULL), 1);
    mbar_init(bar(A2_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =========================== bulk-copy producer ===========================
    if (lane == 0 && NU > 0) {
      mbar_expect_tx(bar(W2_FULL), kW2Bytes);
      bulk_g2s(base + kOffW2, p.packed, kW2Bytes, bar(W2_FULL));
      const char* w3 = p.packed + packed_off_w3();
      uint32_t n = 0;
      for (int u = 0; u < NU; ++u) {
        for (int s = 0; s < NSLAB; ++s, ++n) {
          const uint32_t stage = n % kStages, use = n / kStages;
          mbar_wait(bar(W3_EMPTY + stage), (use & 1u) ^ 1u);
          mbar_expect_tx(bar(W3_FULL + stage), kSlabBytes);
          const uint32_t dst = base + kOffW3 + stage * kSlabBytes;
          const char* src = w3 + (size_t)s * kSlabBytes;
#pragma unroll
          for (int c = 0; c < 4; ++c) bulk_g2s(dst + c * 8192u, src + c * 8192, 8192u, bar(W3_FULL + stage));
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (converged warp, one elected lane issues) ===========================
    if (NU > 0) {
      long long dw[5] = {0, 0, 0, 0, 0};
      const long long t_role = clock64();
      mbar_wait(bar(W2_FULL), 0);
      uint32_t j = 0, slab_n = 0;
      auto issue_layer2 = [&](int u) {  // one job: both halves into the two column halves of slot j&1
        const uint32_t uph = (uint32_t)u & 1u, slot = j & 1u;
        mbar_wait_t(bar(H1_FULL + 0), uph, dw[0]);
        mbar_wait_t(bar(H1_FULL + 1), uph, dw[0]);
        mbar_wait_t(bar(ACC_EMPTY + slot), ((j >> 1) & 1u) ^ 1u, dw[1]);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t w2_desc = umma_desc_sw128(base + kOffW2);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t a_desc = umma_desc_sw128(base + kOffH1 + h * kKBlockBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma(tmem + 256u * slot + 128u * h, a_desc + 2u * k, w2_desc + 2u * k, kIdescN128, k > 0);
          }
          umma_commit(bar(A2_FULL));
          umma_commit(bar(H1_EMPTY + 0));
          umma_commit(bar(H1_EMPTY + 1));
        }
        __syncwarp();
        ++j;
      };
      issue_layer2(0);
      for (int u = 0; u < NU; ++u) {
        const uint32_t uph = (uint32_t)u & 1u;
        for (int s = 0; s < NSLAB; ++s, ++slab_n) {
          if (s == kL2At && u + 1 < NU) issue_layer2(u + 1);
          const uint32_t stage = slab_n % kStages, slot = j & 1u;
          mbar_wait_t(bar(W3_FULL + stage), (slab_n / kStages) & 1u, dw[2]);
          if (s == 0) {
            mbar_wait_t(bar(H2_FULL + 0), uph, dw[3]);
            mbar_wait_t(bar(H2_FULL + 1), uph, dw[3]);
          }
          mbar_wait_t(bar(ACC_EMPTY + slot), ((j >> 1) & 1u) ^ 1u, dw[4]);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a0 = umma_desc_sw128(base + kOffW3 + stage * kSlabBytes);
            const uint64_t b0 = umma_desc_sw128(base + kOffH2);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint32_t ka = (uint32_t)(kk >> 2) * (kKBlockBytes >> 4) + (uint32_t)(kk & 3) * 2u;
              const uint32_t kb = (uint32_t)(kk >> 2) * (kH2KBlockBytes >> 4) + (uint32_t)(kk & 3) * 2u;
              umma(tmem + 256u * slot, a0 + ka, b0 + kb, kIdescN256, kk > 0);
            }
            umma_commit(bar(S_FULL + slot));
            umma_commit(bar(W3_EMPTY + stage));
            if (s == NSLAB - 1) umma_commit(bar(H2_EMPTY));
          }
          __syncwarp();
          ++j;
        }
      }
      // wait until the epilogue has drained the last job (this warp has observed every phase of ACC_EMPTY, so the
      // parity wait is exact): by then every MMA, and every commit queued behind the last one, has retired
      mbar_wait(bar(ACC_EMPTY + ((j - 1u) & 1u)), ((j - 1u) >> 1) & 1u);
      if (p.dbg && lane == 0) {
        for (int i = 0; i < 5; ++i) p.dbg[blockIdx.x * 32 + i] = dw[i];
        p.dbg[blockIdx.x * 32 + 5] = clock64() - t_role;
      }
    }
  } else if (warp >= 4 && warp < 8) {
   if constexpr (TRAIN) {
    // ============ layer-3 epilogue, train mode:
    // running max + arg-max point per (cloud, channel) and the per-channel sum / sum of squares over all real
    // points (BN3 batch statistics).  The arg-max rides in the low byte of the value: each accumulator word gets
    // its position inside the 16-column chunk in place of its 8 lowest mantissa bits (one PRMT), then the usual
    // 3-input max tree runs on those words; one compare per chunk tracks the chunk.  No per-value compare/select
    // chain.  The returned maximum is exact to 2^-15 relative.
    const int q = warp & 3;
    float rmax[NSLAB], rs[NSLAB], rq[NSLAB];
    int ridx[NSLAB];
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) { rmax[s] = -INFINITY; rs[s] = 0.f; rq[s] = 0.f; ridx[s] = 0; }
    int cur_cloud = -1;
    uint32_t seen0 = 0, seen1 = 0;
    auto flush = [&](int cloud) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const float v = __uint_as_float(__float_as_uint(rmax[s]) & 0xffffff00u);
        const unsigned long long key = ((unsigned long long)enc_ordered(v) << 32) |
                                       (unsigned long long)(0xffffffffu - (uint32_t)ridx[s]);
        atomicMax(p.enc64 + (size_t)cloud * p.F + s * 128 + q * 32 + lane, key);
        rmax[s] = -INFINITY;
      }
    };
    for (int u = 0; u < NU; ++u) {
      const int unit = u_begin + u;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnitPts;
      const int ndup = p_base + kUnitPts > p.P ? p_base + kUnitPts - p.P : 0;  // padded columns repeat the last point
      if (cloud != cur_cloud) {
        if (cur_cloud >= 0) flush(cur_cloud);
        cur_cloud = cloud;
      }
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const uint32_t j = job_s(u, s), slot = j & 1u;
        const uint32_t seen = slot ? seen1 : seen0;
        mbar_wait(bar(S_FULL + slot), seen & 1u);
        if (slot) ++seen1; else ++seen0;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 256u * slot;
        float m = -INFINITY;
        int mc = 0;
        unsigned long long sa = 0ull, sb = 0ull, qa = 0ull, qb = 0ull;
        uint32_t ra[16], rb[16];
        // one 16-column chunk: position inside the chunk -> low byte, 3-input max tree, one compare per chunk
        auto chunk = [&](const uint32_t* cur, int c) {
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = __uint_as_float(__byte_perm(cur[i], (uint32_t)i, 0x3214));
          const float t0 = max3(w[0], w[1], w[2]), t1 = max3(w[3], w[4], w[5]), t2 = max3(w[6], w[7], w[8]);
          const float t3 = max3(w[9], w[10], w[11]), t4 = max3(w[12], w[13], w[14]);
          const float cm = fmaxf(max3(t0, t1, t2), max3(t3, t4, w[15]));
          if (cm > m) { m = cm; mc = c; }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            add2(sa, cur[i], cur[i + 1]);     add2(sb, cur[i + 2], cur[i + 3]);
            sq2(qa, cur[i], cur[i + 1]);      sq2(qb, cur[i + 2], cur[i + 3]);
          }
        };
        tmem_ld16(taddr, ra);
        tmem_ld_wait();
        // rolled (the fully unrolled epilogue was instruction-fetch bound: 136 KB of straight-line code per unit)
#pragma unroll 1
        for (int c = 0; c < 16; c += 2) {
          tmem_ld16(taddr + 16u * (uint32_t)(c + 1), rb);
          chunk(ra, c);
          tmem_ld_wait();
          if (c + 2 < 16) tmem_ld16(taddr + 16u * (uint32_t)(c + 2), ra);
          chunk(rb, c + 1);
          tmem_ld_wait();
        }
        const float ylast = __uint_as_float(rb[15]);  // column 255: the last real point whenever ndup > 0
        tc_fence_before();
        mbar_arrive_n(bar(ACC_EMPTY + slot), 2u);  // 128 epilogue threads stand in for the slot's 256 arrivals
        rs[s] += pair_sum(sa) + pair_sum(sb) - (float)ndup * ylast;
        rq[s] += pair_sum(qa) + pair_sum(qb) - (float)ndup * ylast * ylast;
        if (m > rmax[s]) {
          rmax[s] = m;
          const int pt = p_base + 16 * mc + (int)(__float_as_uint(m) & 0xfu);
          ridx[s] = pt < p.P ? pt : p.P - 1;
        }
      }
    }
    if (cur_cloud >= 0) flush(cur_cloud);
    if (NU > 0) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        atomicAdd(p.sum3 + s * 128 + q * 32 + lane, (double)rs[s]);
        atomicAdd(p.sq3 + s * 128 + q * 32 + lane, (double)rq[s]);
      }
    }
   } else {
    // =========================== layer-3 epilogue: running max over the 256 points of a job ===================
    const int q = warp & 3;
    float rmax[NSLAB];
#pragma unroll
    for (int s = 0; s < NSLAB; ++s) rmax[s] = -INFINITY;
    int cur_cloud = -1;
    uint32_t seen0 = 0, seen1 = 0;
    long long dwait = 0;
    const long long t_role = clock64();
    auto flush = [&](int cloud) {
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        atomicMax(p.enc + (size_t)cloud * p.F + s * 128 + q * 32 + lane, enc_ordered(rmax[s]));
        rmax[s] = -INFINITY;
      }
    };
    for (int u = 0; u < NU; ++u) {
      const int cloud = (u_begin + u) / p.tiles_per_cloud;
      if (cloud != cur_cloud) {
        if (cur_cloud >= 0) flush(cur_cloud);
        cur_cloud = cloud;
      }
#pragma unroll
      for (int s = 0; s < NSLAB; ++s) {
        const uint32_t j = job_s(u, s), slot = j & 1u;
        const uint32_t seen = slot ? seen1 : seen0;  // layer-3 jobs already drained from this slot
        mbar_wait_t(bar(S_FULL + slot), seen & 1u, dwait);
        if (slot) ++seen1; else ++seen0;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 256u * slot;
        float m0 = rmax[s], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        // software pipeline over 8 chunks of 32 columns: the load of chunk c+1 is in flight while chunk c is reduced
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr, ra);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t* cur = (c & 1) ? rb : ra;
          uint32_t* nxt = (c & 1) ? ra : rb;
          if (c + 1 < 8) tmem_ld32(taddr + 32u * (c + 1), nxt);
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            m0 = max3(m0, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
            m1 = max3(m1, __uint_as_float(cur[8 + i]), __uint_as_float(cur[9 + i]));
            m2 = max3(m2, __uint_as_float(cur[16 + i]), __uint_as_float(cur[17 + i]));
            m3 = max3(m3, __uint_as_float(cur[24 + i]), __uint_as_float(cur[25 + i]));
          }
          if (c + 1 < 8) tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive_n(bar(ACC_EMPTY + slot), 2u);  // 128 epilogue threads stand in for the slot's 256 arrivals
        rmax[s] = fmaxf(max3(m0, m1, m2), m3);
      }
    }
    if (cur_cloud >= 0) flush(cur_cloud);
    if (p.dbg && q == 0 && lane == 0) {
      p.dbg[blockIdx.x * 32 + 8] = dwait;
      p.dbg[blockIdx.x * 32 + 9] = clock64() - t_role;
    }
   }
  } else if (warp >= 8) {
    // =========================== front end: layer 1 + layer-2 epilogue ======================================
    // two groups of 4 warps, group g owns half g of every unit
    constexpr int NG = 1;
    const int g0 = warp >= 12 ? 1 : 0;
    const int t = threadIdx.x & 127;  // point row inside the half; also the TMEM lane
    const int q = warp & 3;
    const float4* w1p = reinterpret_cast<const float4*>(sm + kOffPar);
    const float4* b2f = reinterpret_cast<const float4*>(sm + kOffPar + 64 * 16);
    long long dw[6] = {0, 0, 0, 0, 0, 0};
    const long long t_role = clock64();

    auto layer1 = [&](int u) {
      const int unit = u_begin + u;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnitPts;
      const uint32_t uph = (uint32_t)u & 1u;
      const float* xc = p.x + (size_t)cloud * 3 * p.P;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const int g = g0 + gi;
        int pt = p_base + g * kHalfPts + t;
        pt = pt < p.P ? pt : p.P - 1;  // ragged tail: repeat the last real point (max is idempotent)
        const float x0 = __ldg(xc + pt), x1 = __ldg(xc + p.P + pt), x2 = __ldg(xc + 2 * p.P + pt);
        mbar_wait_t(bar(H1_EMPTY + g), uph ^ 1u, dw[0]);
        const long long t_l1 = clock64();
        uint8_t* dst = sm + kOffH1 + g * kKBlockBytes;
#pragma unroll
        for (int cg = 0; cg < 8; ++cg) {
          float v[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 w = w1p[cg * 8 + jj];
            v[jj] = fmaf(w.x, x0, fmaf(w.y, x1, fmaf(w.z, x2, w.w)));
          }
          uint4 o;
          o.x = pack_relu_bf16(v[0], v[1]); o.y = pack_relu_bf16(v[2], v[3]);
          o.z = pack_relu_bf16(v[4], v[5]); o.w = pack_relu_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(dst + sw128_off(t, cg * 8)) = o;
        }
        fence_proxy_async();
        mbar_arrive(bar(H1_FULL + g));
        dw[4] += clock64() - t_l1;
      }
    };

    if (NU > 0) layer1(0);
    for (int u = 0; u < NU; ++u) {
      const uint32_t uph = (uint32_t)u & 1u;
      const uint32_t j = job_l2(u), slot = j & 1u;
      mbar_wait_t(bar(A2_FULL), uph, dw[1]);
      tc_fence_after();
      const long long t_e2 = clock64();
      uint4 pk[NG][16];  // this point's 128 layer-2 channels per half: bias + ReLU + bf16, 8 channels per 16 bytes
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + 256u * slot + 128u * (g0 + gi);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + 32u * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int ch = c * 32 + g8 * 8;
            const float4 ba = b2f[ch / 4], bb = b2f[ch / 4 + 1];
            uint4 o;
            if constexpr (TRAIN) {  // BN2 with batch statistics: z2 = sc2 * (raw W2 . h1) + sh2
              const float4 sa = b2f[32 + ch / 4], sb = b2f[32 + ch / 4 + 1];
              o.x = pack_relu_bf16(fmaf(__uint_as_float(r[g8 * 8 + 0]), sa.x, ba.x), fmaf(__uint_as_float(r[g8 * 8 + 1]), sa.y, ba.y));
              o.y = pack_relu_bf16(fmaf(__uint_as_float(r[g8 * 8 + 2]), sa.z, ba.z), fmaf(__uint_as_float(r[g8 * 8 + 3]), sa.w, ba.w));
              o.z = pack_relu_bf16(fmaf(__uint_as_float(r[g8 * 8 + 4]), sb.x, bb.x), fmaf(__uint_as_float(r[g8 * 8 + 5]), sb.y, bb.y));
              o.w = pack_relu_bf16(fmaf(__uint_as_float(r[g8 * 8 + 6]), sb.z, bb.z), fmaf(__uint_as_float(r[g8 * 8 + 7]), sb.w, bb.w));
            } else {
              o.x = pack_relu_bf16(__uint_as_float(r[g8 * 8 + 0]) + ba.x, __uint_as_float(r[g8 * 8 + 1]) + ba.y);
              o.y = pack_relu_bf16(__uint_as_float(r[g8 * 8 + 2]) + ba.z, __uint_as_float(r[g8 * 8 + 3]) + ba.w);
              o.z = pack_relu_bf16(__uint_as_float(r[g8 * 8 + 4]) + bb.x, __uint_as_float(r[g8 * 8 + 5]) + bb.y);
              o.w = pack_relu_bf16(__uint_as_float(r[g8 * 8 + 6]) + bb.z, __uint_as_float(r[g8 * 8 + 7]) + bb.w);
            }
            pk[gi][c * 4 + g8] = o;
          }
        }
        tc_fence_before();
        mbar_arrive(bar(ACC_EMPTY + slot));   // this half of the slot is free again; the values live in registers now
      }
      dw[5] += clock64() - t_e2;
      mbar_wait_t(bar(H2_EMPTY), uph ^ 1u, dw[2]);  // layer 3 of the previous unit has finished reading h2
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const int g = g0 + gi;
        uint8_t* dst = sm + kOffH2;
        const int row = g * kHalfPts + t;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ch = i * 8;
          *reinterpret_cast<uint4*>(dst + (ch >> 6) * kH2KBlockBytes + sw128_off(row, ch & 63)) = pk[gi][i];
        }
        fence_proxy_async();
        mbar_arrive(bar(H2_FULL + g));
      }
      if constexpr (TRAIN) {  // keep these tiles of h2 for backward, in the same swizzled operand image
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          char* gt = p.h2img + ((size_t)(u_begin + u) * 2 + g0 + gi) * kTileBytes;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ch = i * 8;
            *reinterpret_cast<uint4*>(gt + (ch >> 6) * kKBlockBytes + sw128_off(t, ch & 63)) = pk[gi][i];
          }
        }
      }
      if (u + 1 < NU) layer1(u + 1);  // overlaps with layer 3 of unit u on the tensor pipe
    }
    if (p.dbg && t == 0 && g0 == 0) {
      p.dbg[blockIdx.x * 32 + 12] = dw[0]; p.dbg[blockIdx.x * 32 + 13] = dw[1]; p.dbg[blockIdx.x * 32 + 14] = dw[2];
      p.dbg[blockIdx.x * 32 + 15] = clock64() - t_role; p.dbg[blockIdx.x * 32 + 16] = dw[4]; p.dbg[blockIdx.x * 32 + 17] = dw[5];
      p.dbg[blockIdx.x * 32 + 18] = NU;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// out[b,c] = max over points (decoded) + folded BN3 shift
__global__ void __launch_bounds__(256) pointnet_finalize_kernel(const uint32_t* __restrict__ enc,
                                                                const float* __restrict__ shift3, int B, int F,
                                                                float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * F) out[i] = dec_ordered(enc[i]) + shift3[i % F];
}

struct PackParams {
  const float *c1w, *c1b, *c2w, *c2b, *c3w, *c3b;
  const float *g1, *be1, *m1, *v1, *g2, *be2, *m2, *v2, *g3, *be3, *m3, *v3;
  float eps;
  int F;
  char* packed;
};

// Fold eval-mode BN into the convolutions (fp32) and write the bf16 operand images + fp32 side parameters.
__global__ void __launch_bounds__(256) pointnet_pack_kernel(const PackParams a) {
  const int F = a.F;
  const int nW2 = 128 * 64, nW3 = F * 128, nW1 = 64, nB2 = 128, nS3 = F;
  const int total = nW2 + nW3 + nW1 + nB2 + nS3;
  __nv_bfloat16* w2img = reinterpret_cast<__nv_bfloat16*>(a.packed);
  char* w3img = a.packed + packed_off_w3();
  float* par = reinterpret_cast<float*>(a.packed + packed_off_par(F));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < nW2) {
      const int c = i / 64, k = i % 64;
      const float sc = a.g2[c] * rsqrtf(a.v2[c] + a.eps);
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(w2img) + sw128_off(c, k)) =
          __float2bfloat16_rn(a.c2w[c * 64 + k] * sc);
    } else if (i < nW2 + nW3) {
      const int j = i - nW2;
      const int c = j / 128, k = j % 128;
      const float sc = a.g3[c] * rsqrtf(a.v3[c] + a.eps);
      char* slab = w3img + (size_t)(c / 128) * kSlabBytes + (size_t)(k / 64) * kKBlockBytes;
      *reinterpret_cast<__nv_bfloat16*>(slab + sw128_off(c % 128, k % 64)) = __float2bfloat16_rn(a.c3w[c * 128 + k] * sc);
    } else if (i < nW2 + nW3 + nW1) {
      const int c = i - nW2 - nW3;
      const float sc = a.g1[c] * rsqrtf(a.v1[c] + a.eps);
      par[c * 4 + 0] = a.c1w[c * 3 + 0] * sc;
      par[c * 4 + 1] = a.c1w[c * 3 + 1] * sc;
      par[c * 4 + 2] = a.c1w[c * 3 + 2] * sc;
      par[c * 4 + 3] = (a.c1b[c] - a.m1[c]) * sc + a.be1[c];
    } else if (i < nW2 + nW3 + nW1 + nB2) {
      const int c = i - nW2 - nW3 - nW1;
      const float sc = a.g2[c] * rsqrtf(a.v2[c] + a.eps);
      par[64 * 4 + c] = (a.c2b[c] - a.m2[c]) * sc + a.be2[c];
    } else {
      const int c = i - nW2 - nW3 - nW1 - nB2;
      const float sc = a.g3[c] * rsqrtf(a.v3[c] + a.eps);
      par[64 * 4 + 128 + c] = (a.c3b[c] - a.m3[c]) * sc + a.be3[c];
    }
  }
}

}  // namespace pn
}  // namespace crdpn

using namespace crdpn;

using crdpn::pn::pointnet_f_ok;

extern "C" int crdpn_pointnet_packed_bytes(int64_t F, size_t* bytes) {
  if (!bytes) return fail(CRDPN_E_BADARG, "crdpn_pointnet_packed_bytes: null pointer");
  if (!pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  *bytes = pn::packed_bytes((int)F);
  return CRDPN_OK;
}

extern "C" int crdpn_pointnet_pack(const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                                   const float* conv3_w, const float* conv3_b,
                                   const float* bn1_w, const float* bn1_b, const float* bn1_mean, const float* bn1_var,
                                   const float* bn2_w, const float* bn2_b, const float* bn2_mean, const float* bn2_var,
                                   const float* bn3_w, const float* bn3_b, const float* bn3_mean, const float* bn3_var,
                                   float bn_eps, int64_t F, void* packed, void* stream) {
  if (!conv1_w || !conv1_b || !conv2_w || !conv2_b || !conv3_w || !conv3_b || !bn1_w || !bn1_b || !bn1_mean || !bn1_var ||
      !bn2_w || !bn2_b || !bn2_mean || !bn2_var || !bn3_w || !bn3_b || !bn3_mean || !bn3_var || !packed)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_pack: null pointer");
  if (!pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  if ((uintptr_t)packed & 15) return fail(CRDPN_E_ALIGN, "crdpn_pointnet_pack: packed buffer must be 16-byte aligned");
  pn::PackParams a{conv1_w, conv1_b, conv2_w, conv2_b, conv3_w, conv3_b, bn1_w, bn1_b, bn1_mean, bn1_var,
                   bn2_w, bn2_b, bn2_mean, bn2_var, bn3_w, bn3_b, bn3_mean, bn3_var, bn_eps, (int)F, (char*)packed};
  pn::pointnet_pack_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(a);
  CRDPN_LAUNCH_CHECK("pointnet_pack_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_pointnet_workspace_bytes(int64_t B, int64_t P, int64_t F, int device, size_t* bytes) {
  (void)device;
  if (!bytes || B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_workspace_bytes: bad argument");
  if (!pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  *bytes = (size_t)B * (size_t)F * 4 + pn::kDbgBytes;
  return CRDPN_OK;
}

template <int NSLAB>
static int launch_pointnet(const pn::FwdParams& fp, int grid, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  if (!attr_set[device]) {
    CRDPN_CUDA(cudaFuncSetAttribute(pn::pointnet_fwd_kernel_v2<NSLAB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)pn::kSmemAlloc));
    CRDPN_CUDA(cudaFuncSetAttribute(pn::pointnet_fwd_kernel_v2<NSLAB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)pn::kSmemAlloc));
    attr_set[device] = true;
  }
  {
    ScopedKernelTimer tm(CRDPN_K_POINTNET_FWD, st);
    if (fp.train_par) pn::pointnet_fwd_kernel_v2<NSLAB, true><<<grid, pn::kThreads, pn::kSmemAlloc, st>>>(fp);  // train
    else pn::pointnet_fwd_kernel_v2<NSLAB, false><<<grid, pn::kThreads, pn::kSmemAlloc, st>>>(fp);                // v2: N=256
  }
  CRDPN_LAUNCH_CHECK("pointnet_fwd_kernel_v2");
  return CRDPN_OK;
}

int crdpn::pn::launch_fwd(const FwdParams& fp, int grid, cudaStream_t st) {
  switch (fp.F / 128) {
    case 1: return launch_pointnet<1>(fp, grid, st);
    case 2: return launch_pointnet<2>(fp, grid, st);
    case 4: return launch_pointnet<4>(fp, grid, st);
    default: return launch_pointnet<8>(fp, grid, st);
  }
}

extern "C" int crdpn_pointnet_forward_eval(const float* x, int64_t B, int64_t P, int64_t F, const void* packed,
                                           float* out, void* workspace, size_t workspace_bytes, int variant,
                                           void* stream) {
  if (!x || !packed || !out || !workspace) return fail(CRDPN_E_BADARG, "crdpn_pointnet_forward_eval: null pointer");
  if (B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_forward_eval: bad size");
  if (!pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  if (workspace_bytes < (size_t)B * (size_t)F * 4 + ((variant & 4) ? pn::kDbgBytes : 0)) return fail(CRDPN_E_WORKSPACE, "crdpn_pointnet_forward_eval: workspace too small");
  if (((uintptr_t)packed & 15) || ((uintptr_t)workspace & 3)) return fail(CRDPN_E_ALIGN, "crdpn_pointnet_forward_eval: alignment");
  if (B * ((P + pn::kUnitPts - 1) / pn::kUnitPts) >= (1ll << 30)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_forward_eval: too many tiles");
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  if (di.max_smem_optin < (int)pn::kSmemAlloc) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_forward_eval: not enough shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  pn::FwdParams fp;
  fp.x = x; fp.B = (int)B; fp.P = (int)P; fp.F = (int)F;
  fp.packed = (const char*)packed;
  fp.enc = (uint32_t*)workspace;
  fp.tiles_per_cloud = (int)((P + pn::kUnitPts - 1) / pn::kUnitPts);
  fp.total_units = (int)B * fp.tiles_per_cloud;
  fp.flags = variant;
  fp.train_par = nullptr; fp.h2img = nullptr; fp.enc64 = nullptr; fp.sum3 = nullptr; fp.sq3 = nullptr;
  fp.dbg = (variant & 4) ? (long long*)((char*)workspace + (((size_t)B * (size_t)F * 4 + 15) & ~(size_t)15)) : nullptr;
  CRDPN_CUDA(cudaMemsetAsync(workspace, 0, (size_t)B * (size_t)F * 4, st));
  const int grid = fp.total_units < di.sms ? fp.total_units : di.sms;
  rc = pn::launch_fwd(fp, grid, st);
  if (rc) return rc;
  const float* shift3 = reinterpret_cast<const float*>((const char*)packed + pn::packed_off_par((int)F) + 64 * 16 + 128 * 4);
  const int n = (int)(B * F);
  pn::pointnet_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>((const uint32_t*)workspace, shift3, (int)B, (int)F, out);
  CRDPN_LAUNCH_CHECK("pointnet_finalize_kernel");
  return CRDPN_OK;
}
