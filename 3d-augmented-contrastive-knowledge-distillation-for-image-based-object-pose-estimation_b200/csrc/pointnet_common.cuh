// pointnet_common.cuh -- constants and sm_100a PTX helpers (mbarrier, bulk copy, tcgen05, TMEM) shared by the
// PointNet encoder kernels (pointnet_kernels.cu: eval + train forward; pointnet_train.cu: statistics + backward).
#pragma once
#include "common.cuh"

namespace crdpn {
namespace pn {

constexpr int kThreads = 512;
constexpr int kHalfPts = 128;
constexpr int kUnitPts = 256;
constexpr int kStages = 3;
constexpr uint32_t kSlabBytes = 32768;   // 128 rows x 128 k bf16 = two 16 KB K-blocks
constexpr uint32_t kKBlockBytes = 16384; // 128 rows x 64 k bf16 (one 128-byte swizzle span per row)
constexpr uint32_t kW2Bytes = 16384;
constexpr size_t kDbgBytes = 256 * 32 * 8;  // optional per-CTA cycle counters behind the max buffer

// shared memory map (offsets from a 1024-aligned base)
constexpr uint32_t kOffW3 = 0;
constexpr uint32_t kOffH2 = kOffW3 + kStages * kSlabBytes;   // 2 halves x 32 KB
constexpr uint32_t kOffH1 = kOffH2 + 2 * kSlabBytes;         // 2 halves x 16 KB
constexpr uint32_t kOffW2 = kOffH1 + 2 * kKBlockBytes;
constexpr uint32_t kOffPar = kOffW2 + kW2Bytes;              // W1p[64][4] f32, b2f[128] f32, (train: sc2[128] f32)
constexpr uint32_t kParBytes = 64 * 16 + 128 * 4 + 128 * 4;
constexpr uint32_t kTileBytes = 32768;  // one 128-point tile of h2 (two 64-channel K-blocks) in the train context
constexpr uint32_t kOffBar = kOffPar + kParBytes;
constexpr uint32_t kNumBars = 32;
constexpr uint32_t kSmemBytes = kOffBar + kNumBars * 8 + 16;
constexpr uint32_t kSmemAlloc = kSmemBytes + 1024;           // slack for manual 1024-byte alignment


// packed parameter buffer (global), produced by pointnet_pack_kernel
__host__ __device__ inline size_t packed_off_w3() { return kW2Bytes; }
__host__ __device__ inline size_t packed_off_par(int F) { return kW2Bytes + (size_t)(F / 128) * kSlabBytes; }
__host__ __device__ inline size_t packed_bytes(int F) { return packed_off_par(F) + kParBytes + (size_t)F * 4; }

// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile of 64 bf16 per row
__host__ __device__ inline uint32_t sw128_off(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

// ---------------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ unsigned int g_pointnet_timeout = 0;  // set when a wait gives up (kernel then traps instead of hanging)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) {  // ~2 s: a protocol bug, never a legitimate wait
      atomicExch(&g_pointnet_timeout, 1u + (bar & 0xffffu));
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
// packed pair arithmetic (sm_100 FADD2 / FFMA2): acc.{lo,hi} += {a,b};  acc.{lo,hi} += {a*a, b*b}
__device__ __forceinline__ void add2(unsigned long long& acc, uint32_t a, uint32_t b) {
  asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %2};\n\tadd.rn.f32x2 %0, %0, t;\n\t}" : "+l"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ void sq2(unsigned long long& acc, uint32_t a, uint32_t b) {
  asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %2};\n\tfma.rn.f32x2 %0, t, t, %0;\n\t}" : "+l"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ float pair_sum(unsigned long long v) {
  return __uint_as_float((uint32_t)v) + __uint_as_float((uint32_t)(v >> 32));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// exactly one lane of a converged warp (lets ptxas issue the uniform-datapath tcgen05 ops without a per-lane loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// relu + round-to-nearest-even bf16 pack: low half <- a, high half <- b
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t enc_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// generic tcgen05.mma (kind::f16: bf16 x bf16 -> fp32), both operands K-major, cta_group::1
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc(volatile uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"(smem_u32((const void*)slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {  // low half <- a, high half <- b (round to nearest even)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

struct FwdParams {
  const float* x;        // [B,3,P]
  int B, P, F;
  const char* packed;
  uint32_t* enc;         // [B,F] order-preserving encoding of the running max (zeroed before launch)
  int tiles_per_cloud;
  int total_units;
  long long* dbg;         // optional [grid][32] cycle counters (flags bit2), else null
  int flags;             // bit0: rotate the slab order per CTA; bit1 (diagnostic, wrong results): load each ring stage once
  // train mode (batch-statistics BatchNorm; pointnet_fwd_kernel_v2<NSLAB, true>)
  const float* train_par;        // [512] f32: W1p[64][4] (BN1 batch statistics folded), sh2[128], sc2[128]
  char* h2img;                   // [total_units*2][32 KB] bf16 tiles of h2 (the shared-memory operand image), kept for backward
  unsigned long long* enc64;     // [B,F] (ordered max << 32) | (0xffffffff - point index)   (zeroed before launch)
  double* sum3;                  // [F] sum over all real points of the sign-folded conv3 output (zeroed)
  double* sq3;                   // [F] sum of squares
};

// launches pointnet_fwd_kernel_v2<F/128, train_par != nullptr> (defined in pointnet_kernels.cu)
int launch_fwd(const FwdParams& fp, int grid, cudaStream_t st);
inline bool pointnet_f_ok(int64_t F) { return F == 128 || F == 256 || F == 512 || F == 1024; }

}  // namespace pn
}  // namespace crdpn
