// umma_tf32_probe.cu -- hardware check of the tensor-core building block the bank-streaming CRD step needs (DESIGN.md
// section 8): tcgen05.mma kind::tf32 on fp32 tiles stored ONCE in shared memory in the K-major SWIZZLE_128B image and read
//   (1) K-major   as A = rows [M = 128 (64 bank-1 rows | 64 bank-2 rows)] x K = 128 features   -> scores  S = A . [V2 | V1]^T
//   (2) MN-major  as A' = features [M = 128] x K = 64 rows of ONE bank (same bytes, descriptor major bit) -> G^T = A' . C
// with B operands K-major ([V2 | V1] : 96 x 128;  C^T : 48 anchors x 64 rows).  One CTA, one tile; the result
// (128 lanes x 192 TMEM columns: scores | G2^T | G1^T) is dumped to global memory for comparison with numpy.
// FINDING (round 1): (1) works as written.  (2) does NOT work on the K-major SWIZZLE_128B image: for 32-bit operands the
// hardware accepts MN-major only in the SWIZZLE_128B_BASE32B layout (descriptor layout type 1: atoms of 4 k-rows x 128 B,
// 32-byte chunks XOR-ed with the k-row index, address bits [5,7) ^= [7,9)); with layout type 2 the MMA silently yields
// zeros.  The probe therefore keeps a second image of each bank's rows in that layout for the gradient GEMM -- the
// streaming kernel will have to write every tile twice (or convert the gradient operand to 16-bit).
// Not a product path: a probe that pins the descriptor encodings before the streaming kernel is built on them.
#include "common.cuh"
#include "../../include/crdpn_b200_dev.h"
#include "pointnet_common.cuh"

namespace crdpn {
namespace probe {

// byte offset of fp32 element (row, k) inside one K-major SWIZZLE_128B K-block (32 floats = 128 bytes per row)
__host__ __device__ inline uint32_t sw128_off_f32(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + (k & 3) * 4);
}

// kind::tf32 instruction descriptor: D = f32 (bit 4), A = B = TF32 (format 2 at bits 7 / 10), major bits 15 / 16
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// shared-memory descriptor, SWIZZLE_128B: start >> 4 | LBO (16-byte units) << 16 | SBO (16-byte units) << 32 | version 1 | layout 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2u) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29);
  return ((uint64_t)hi << 32) | lo;
}
// byte offset of fp32 element (k-row r, feature e) inside one bank's MN-major SWIZZLE_128B_BASE32B image:
// [feature block e/32][r][128 B], 32-byte chunks XOR-ed with r & 3
__host__ __device__ inline uint32_t mn32b_off_f32(int r, int e, int rows) {
  return (uint32_t)((e >> 5) * (rows * 128) + r * 128 + (((((e & 31) >> 3) ^ (r & 3)) & 3) << 5) + (e & 7) * 4);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr uint32_t kA = 0;                    // rows tile: 4 K-blocks x [128 rows x 128 B] = 64 KB
constexpr uint32_t kBV = 65536;               // [V2 | V1]: 4 K-blocks x [96 x 128 B] = 48 KB
constexpr uint32_t kBC2 = kBV + 49152;        // C2^T: 2 K-blocks x [48 x 128 B] = 12 KB
constexpr uint32_t kBC1 = kBC2 + 12288;
constexpr uint32_t kAT1 = kBC1 + 12288;        // bank-1 rows, MN-major BASE32B image: 4 feature blocks x [64 rows x 128 B] = 32 KB
constexpr uint32_t kAT2 = kAT1 + 32768;
constexpr uint32_t kBar = kAT2 + 32768;
constexpr uint32_t kSmem = kBar + 64 + 1024;

__global__ void __launch_bounds__(128, 1) umma_tf32_probe_kernel(const float* __restrict__ rows1, const float* __restrict__ rows2,
                                                                 const float* __restrict__ v1, const float* __restrict__ v2,
                                                                 const float* __restrict__ c1, const float* __restrict__ c2,
                                                                 float* __restrict__ out, int mode) {
  using namespace pn;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar = base + kBar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kBar + 32);
  const int tid = threadIdx.x, warp = tid >> 5;

  // operands into their swizzled images (generic-proxy stores, then fence.proxy.async)
  for (int i = tid; i < 128 * 128; i += 128) {          // A: row m (m < 64: bank 1, else bank 2), feature e
    const int m = i >> 7, e = i & 127;
    const float v = m < 64 ? rows1[m * 128 + e] : rows2[(m - 64) * 128 + e];
    *reinterpret_cast<float*>(sm + kA + (e >> 5) * (128 * 128) + sw128_off_f32(m, e & 31)) = v;
  }
  for (int i = tid; i < 64 * 128; i += 128) {           // the same rows once more, in the MN-major image of each bank
    const int r = i >> 7, e = i & 127;
    *reinterpret_cast<float*>(sm + kAT1 + mn32b_off_f32(r, e, 64)) = rows1[r * 128 + e];
    *reinterpret_cast<float*>(sm + kAT2 + mn32b_off_f32(r, e, 64)) = rows2[r * 128 + e];
  }
  for (int i = tid; i < 96 * 128; i += 128) {           // B_V: row n (n < 48: V2[n], else V1[n - 48]), feature e
    const int n = i >> 7, e = i & 127;
    const float v = n < 48 ? v2[n * 128 + e] : v1[(n - 48) * 128 + e];
    *reinterpret_cast<float*>(sm + kBV + (e >> 5) * (96 * 128) + sw128_off_f32(n, e & 31)) = v;
  }
  for (int i = tid; i < 48 * 64; i += 128) {            // C^T: row b (anchor), k = r (bank row): C[r][b]
    const int b = i >> 6, r = i & 63;
    *reinterpret_cast<float*>(sm + kBC2 + (r >> 5) * (48 * 128) + sw128_off_f32(b, r & 31)) = c2[r * 48 + b];
    *reinterpret_cast<float*>(sm + kBC1 + (r >> 5) * (48 * 128) + sw128_off_f32(b, r & 31)) = c1[r * 48 + b];
  }
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // (1) scores: M = 128, N = 96, K = 128 -> 16 steps of K = 8 (32 bytes inside a 128-byte row; 4 steps per K-block)
      //     mode 1: the same with M = 64 (only the first 64 stacked rows), to record where M = 64 accumulators live in TMEM
      const uint32_t i_s = mode == 1 ? make_idesc_tf32(64, 96, 0, 0) : make_idesc_tf32(128, 96, 0, 0);
      for (int kk = 0; kk < 16; ++kk) {
        const uint64_t da = umma_desc(base + kA + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
        const uint64_t db = umma_desc(base + kBV + (kk >> 2) * (96 * 128) + (kk & 3) * 32, 16, 1024);
        umma_tf32(tmem, da, db, i_s, kk > 0);
      }
      // (2) gradients: A' = a bank's rows read MN-major from the BASE32B image (M = 128 features: 4 blocks of 32 floats,
      //     LBO = one feature block = 64 rows x 128 B; K = rows: 8 per step = two 512-byte atoms of 4 rows, SBO = 512 B),
      //     B = C^T K-major
      constexpr uint32_t i_g = make_idesc_tf32(128, 48, 1, 0);
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da1 = umma_desc(base + kAT1 + kk * 1024, 64 * 128, 512, 1u);         // bank-1 rows 8kk .. 8kk+7
        const uint64_t da2 = umma_desc(base + kAT2 + kk * 1024, 64 * 128, 512, 1u);         // bank-2 rows
        const uint64_t dc2 = umma_desc(base + kBC2 + (kk >> 2) * (48 * 128) + (kk & 3) * 32, 16, 1024);
        const uint64_t dc1 = umma_desc(base + kBC1 + (kk >> 2) * (48 * 128) + (kk & 3) * 32, 16, 1024);
        umma_tf32(tmem + 96, da1, dc2, i_g, kk > 0);    // G2^T[e][b] = sum_r bank1[r][e] C2[r][b]
        umma_tf32(tmem + 144, da2, dc1, i_g, kk > 0);   // G1^T[e][b] = sum_r bank2[r][e] C1[r][b]
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  // dump: thread = TMEM lane (row of the accumulators), 192 columns
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 192; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)tid * 192 + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256u);
  }
}

// ---- mode 2: the same two GEMMs with bf16 operands (kind::f16) from ONE K-major SWIZZLE_128B image of the stacked rows:
// 16-bit operands may be read MN-major from the ordinary SWIZZLE_128B layout, so a single image serves both GEMMs.
constexpr uint32_t kA16 = 0;                   // 2 K-blocks x [128 rows x 128 B] = 32 KB
constexpr uint32_t kBV16 = 32768;              // [V2 | V1]: 2 K-blocks x [96 x 128 B] = 24 KB
constexpr uint32_t kBC216 = kBV16 + 24576;     // C2^T: [48 anchors x 64 rows] bf16 = one K-block, 6 KB
constexpr uint32_t kBC116 = kBC216 + 6144;
constexpr uint32_t kBar16 = kBC116 + 6144;
constexpr uint32_t kSmem16 = kBar16 + 64 + 1024;

__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) umma_bf16_probe_kernel(const float* __restrict__ rows1, const float* __restrict__ rows2,
                                                                 const float* __restrict__ v1, const float* __restrict__ v2,
                                                                 const float* __restrict__ c1, const float* __restrict__ c2,
                                                                 float* __restrict__ out) {
  using namespace pn;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar = base + kBar16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kBar16 + 32);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto put = [&](uint32_t region, int kblock_rows, int row, int k, float v) {   // bf16 element (row, k) of a K-major SW128 image
    *reinterpret_cast<__nv_bfloat16*>(sm + region + (k >> 6) * (kblock_rows * 128) + sw128_off(row, k & 63)) = __float2bfloat16_rn(v);
  };
  for (int i = tid; i < 128 * 128; i += 128) {
    const int m = i >> 7, e = i & 127;
    put(kA16, 128, m, e, m < 64 ? rows1[m * 128 + e] : rows2[(m - 64) * 128 + e]);
  }
  for (int i = tid; i < 96 * 128; i += 128) {
    const int n = i >> 7, e = i & 127;
    put(kBV16, 96, n, e, n < 48 ? v2[n * 128 + e] : v1[(n - 48) * 128 + e]);
  }
  for (int i = tid; i < 48 * 64; i += 128) {
    const int b = i >> 6, r = i & 63;
    put(kBC216, 48, b, r, c2[r * 48 + b]);
    put(kBC116, 48, b, r, c1[r * 48 + b]);
  }
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t i_s = make_idesc_bf16(128, 96, 0, 0);
      for (int kk = 0; kk < 8; ++kk) {          // K = 128 in steps of 16 bf16 = 32 bytes; 4 steps per 64-wide K-block
        const uint64_t da = umma_desc(base + kA16 + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
        const uint64_t db = umma_desc(base + kBV16 + (kk >> 2) * (96 * 128) + (kk & 3) * 32, 16, 1024);
        umma_f16(tmem, da, db, i_s, kk > 0);
      }
      // gradients: the SAME image read MN-major: M = 128 features = 2 blocks of 64 bf16 (LBO = one K-block = 16384 B),
      // K = rows, 16 per step = two 8-row atoms (SBO = 1024 B)
      constexpr uint32_t i_g = make_idesc_bf16(128, 48, 1, 0);
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da1 = umma_desc(base + kA16 + kk * 2048, 16384, 1024);
        const uint64_t da2 = umma_desc(base + kA16 + 64 * 128 + kk * 2048, 16384, 1024);
        const uint64_t dc2 = umma_desc(base + kBC216 + kk * 32, 16, 1024);
        const uint64_t dc1 = umma_desc(base + kBC116 + kk * 32, 16, 1024);
        umma_f16(tmem + 96, da1, dc2, i_g, kk > 0);
        umma_f16(tmem + 144, da2, dc1, i_g, kk > 0);
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 192; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)tid * 192 + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256u);
  }
}

}  // namespace probe
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_umma_tf32_probe(const float* rows1, const float* rows2, const float* v1, const float* v2, const float* c1,
                                     const float* c2, float* out, int mode, void* stream) {
  if (!rows1 || !rows2 || !v1 || !v2 || !c1 || !c2 || !out) return fail(CRDPN_E_BADARG, "crdpn_umma_tf32_probe: null pointer");
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  if (di.max_smem_optin < (int)probe::kSmem) return fail(CRDPN_E_UNSUPPORTED, "crdpn_umma_tf32_probe: not enough shared memory");
  CRDPN_CUDA(cudaFuncSetAttribute(probe::umma_tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe::kSmem));
  CRDPN_CUDA(cudaFuncSetAttribute(probe::umma_bf16_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)probe::kSmem16));
  if (mode == 2) {
    probe::umma_bf16_probe_kernel<<<1, 128, probe::kSmem16, (cudaStream_t)stream>>>(rows1, rows2, v1, v2, c1, c2, out);
    CRDPN_LAUNCH_CHECK("umma_bf16_probe_kernel");
    return CRDPN_OK;
  }
  probe::umma_tf32_probe_kernel<<<1, 128, probe::kSmem, (cudaStream_t)stream>>>(rows1, rows2, v1, v2, c1, c2, out, mode);
  CRDPN_LAUNCH_CHECK("umma_tf32_probe_kernel");
  return CRDPN_OK;
}
