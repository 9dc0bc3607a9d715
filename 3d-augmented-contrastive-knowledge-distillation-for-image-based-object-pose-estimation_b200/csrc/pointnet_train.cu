// pointnet_train.cu -- train-mode (batch-statistics BatchNorm) forward of the PointNet encoder
// (ShapeEncoderPC.forward under model.train(), auxiliary/model.py:174-180, called from training.py:47) and the
// pieces of its backward that live outside the two dense backward kernels (pointnet_backward.cu).
//
// Train-mode BatchNorm needs per-channel statistics over ALL B*P points before the next layer can run, which is
// what blocks the naive single-kernel fusion.  The forward is therefore four steps, none of which materialises
// anything of size B*F*P:
//   1. x moments (sum x, sum x x^T: 9 numbers).  conv1 is affine in x, so BN1's batch mean / variance follow
//      analytically: mean = w.mx + b, var = w^T Cov(x) w.                                  [pn_xmoments, pn_fold1]
//   2. statistics pass for BN2: layer 1 (CUDA cores) + layer 2 on tcgen05 "swap-AB" (accumulator lanes = channels,
//      columns = points), per-thread sum / sum of squares over the columns.  6% of the FLOPs.   [pn_stats2, pn_fold2]
//   3. the fused forward kernel (pointnet_fwd_kernel_v2<NSLAB, true>): BN1/BN2 applied with the batch statistics,
//      layer 3 uses the RAW conv3 weights with only sign(gamma3) folded in, so max over points commutes with BN3;
//      the epilogue keeps, per (cloud, channel), the running max AND its arg-max point, and per channel the
//      sum / sum of squares over all real points (BN3's batch statistics).  h2 is written out once (bf16, in the
//      tensor-core operand image) for backward.
//   4. finalize: BN3 statistics -> out = |gamma3| * (max - mean) * rsqrt(var + eps) + beta3; running statistics of
//      all three BatchNorms are updated in place (momentum, unbiased variance), num_batches_tracked += 1.
// Conv biases cancel inside train-mode BN (they only move the batch mean), so the kernels never add them; they only
// enter the running means.
#include "pointnet_common.cuh"
#include "pointnet_train.cuh"

namespace crdpn {
namespace pn {

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pn_xmoments_kernel(const float* __restrict__ x, int B, int P, double* __restrict__ xmom) {
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long total = (long long)B * P;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < total; n += (long long)gridDim.x * blockDim.x) {
    const long long b = n / P;
    const int pt = (int)(n - b * P);
    const float* xc = x + b * 3 * P;
    const double x0 = xc[pt], x1 = xc[P + pt], x2 = xc[2 * P + pt];
    a[0] += x0; a[1] += x1; a[2] += x2;
    a[3] += x0 * x0; a[4] += x0 * x1; a[5] += x0 * x2; a[6] += x1 * x1; a[7] += x1 * x2; a[8] += x2 * x2;
  }
  __shared__ double red[9][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    double v = a[i];
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(xmom + threadIdx.x, v);
  }
}

// raw bf16 operand images for the train-mode kernels: W2 [128 x 64] and sign(gamma3) * W3 [F x 128] (slabs)
__global__ void __launch_bounds__(256) pn_pack_train_kernel(const float* __restrict__ c2w, const float* __restrict__ c3w,
                                                            const float* __restrict__ g3, int F, char* __restrict__ packed) {
  const int nW2 = 128 * 64, nW3 = F * 128;
  char* w3img = packed + packed_off_w3();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nW2 + nW3; i += gridDim.x * blockDim.x) {
    if (i < nW2) {
      const int c = i / 64, k = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(packed + sw128_off(c, k)) = __float2bfloat16_rn(c2w[i]);
    } else {
      const int j = i - nW2;
      const int c = j / 128, k = j % 128;
      const float sg = g3[c] >= 0.f ? 1.f : -1.f;
      char* slab = w3img + (size_t)(c / 128) * kSlabBytes + (size_t)(k / 64) * kKBlockBytes;
      *reinterpret_cast<__nv_bfloat16*>(slab + sw128_off(c % 128, k % 64)) = __float2bfloat16_rn(c3w[j] * sg);
    }
  }
}

struct Fold1Params {
  const float *c1w, *c1b, *g1, *be1;
  float *rm1, *rv1;
  long long *nbt1, *nbt2, *nbt3;
  const double* xmom;
  double M;
  float eps, momentum;
  float* stats;      // TrainCtx stats block
  float* train_par;  // [512]
  double* xstat;     // [12]: mean x (3), covariance (9), kept for backward
};

__global__ void pn_fold1_kernel(const Fold1Params a) {
  const int c = threadIdx.x;
  const double M = a.M;
  double mx[3], cov[3][3];
  for (int d = 0; d < 3; ++d) mx[d] = a.xmom[d] / M;
  const int tri[3][3] = {{3, 4, 5}, {4, 6, 7}, {5, 7, 8}};
  for (int d = 0; d < 3; ++d)
    for (int e = 0; e < 3; ++e) cov[d][e] = a.xmom[tri[d][e]] / M - mx[d] * mx[e];
  if (c < 64) {
    const double w0 = a.c1w[c * 3 + 0], w1 = a.c1w[c * 3 + 1], w2 = a.c1w[c * 3 + 2];
    const double w[3] = {w0, w1, w2};
    double mean_raw = 0, var = 0;
    for (int d = 0; d < 3; ++d) {
      mean_raw += w[d] * mx[d];
      for (int e = 0; e < 3; ++e) var += w[d] * cov[d][e] * w[e];
    }
    if (var < 0) var = 0;
    const double istd = 1.0 / sqrt(var + (double)a.eps);
    const double sc = (double)a.g1[c] * istd;
    a.train_par[c * 4 + 0] = (float)(w0 * sc);
    a.train_par[c * 4 + 1] = (float)(w1 * sc);
    a.train_par[c * 4 + 2] = (float)(w2 * sc);
    a.train_par[c * 4 + 3] = (float)((double)a.be1[c] - mean_raw * sc);
    const double mean = mean_raw + (double)a.c1b[c];
    a.stats[kStatMean1 + c] = (float)mean_raw;
    a.stats[kStatIstd1 + c] = (float)istd;
    const double mom = a.momentum;
    a.rm1[c] = (float)((1.0 - mom) * (double)a.rm1[c] + mom * mean);
    a.rv1[c] = (float)((1.0 - mom) * (double)a.rv1[c] + mom * var * (M / (M > 1 ? M - 1 : 1)));
  }
  if (c == 0) {
    *a.nbt1 += 1; *a.nbt2 += 1; *a.nbt3 += 1;
    for (int d = 0; d < 3; ++d) {
      a.xstat[d] = mx[d];
      for (int e = 0; e < 3; ++e) a.xstat[3 + d * 3 + e] = cov[d][e];
    }
  }
}

struct Fold2Params {
  const float *c2b, *g2, *be2;
  float *rm2, *rv2;
  const double *sum2, *sq2;
  double M;
  float eps, momentum;
  float* stats;
  float* train_par;
};

__global__ void pn_fold2_kernel(const Fold2Params a) {
  const int c = threadIdx.x;
  if (c >= 128) return;
  const double M = a.M;
  const double mean_raw = a.sum2[c] / M;
  double var = a.sq2[c] / M - mean_raw * mean_raw;
  if (var < 0) var = 0;
  const double istd = 1.0 / sqrt(var + (double)a.eps);
  const double sc = (double)a.g2[c] * istd;
  a.train_par[256 + c] = (float)((double)a.be2[c] - mean_raw * sc);  // sh2
  a.train_par[384 + c] = (float)sc;                                  // sc2
  const double mean = mean_raw + (double)a.c2b[c];
  a.stats[kStatMean2 + c] = (float)mean_raw;
  a.stats[kStatIstd2 + c] = (float)istd;
  const double mom = a.momentum;
  a.rm2[c] = (float)((1.0 - mom) * (double)a.rm2[c] + mom * mean);
  a.rv2[c] = (float)((1.0 - mom) * (double)a.rv2[c] + mom * var * (M / (M > 1 ? M - 1 : 1)));
}

struct FinalizeParams {
  const unsigned long long* enc64;
  const double *sum3, *sq3;
  const float *c3b, *g3, *be3;
  float *rm3, *rv3;
  double M;
  float eps, momentum;
  int B, F;
  float* stats;
  int* argmax;
  float* yhat3;
  float* out;
};

__global__ void __launch_bounds__(256) pn_train_finalize_kernel(const FinalizeParams a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.F) return;
  const int c = i % a.F;
  const double M = a.M;
  const double mean_f = a.sum3[c] / M;  // of the sign-folded, bias-free conv3 output
  double var = a.sq3[c] / M - mean_f * mean_f;
  if (var < 0) var = 0;
  const double istd = 1.0 / sqrt(var + (double)a.eps);
  const unsigned long long key = a.enc64[i];
  const float val = dec_ordered((uint32_t)(key >> 32));
  const uint32_t idx = 0xffffffffu - (uint32_t)key;
  const float g = a.g3[c];
  const float sg = g >= 0.f ? 1.f : -1.f;
  const float yh = (float)(((double)val - mean_f) * istd);
  a.out[i] = fmaf(fabsf(g), yh, a.be3[c]);
  a.yhat3[i] = sg * yh;
  a.argmax[i] = (int)idx;
  if (i < a.F) {
    const double mean = (double)sg * mean_f + (double)a.c3b[c];
    a.stats[kStatMean3 + c] = (float)((double)sg * mean_f);
    a.stats[kStatIstd3(a.F) + c] = (float)istd;
    const double mom = a.momentum;
    a.rm3[c] = (float)((1.0 - mom) * (double)a.rm3[c] + mom * mean);
    a.rv3[c] = (float)((1.0 - mom) * (double)a.rv3[c] + mom * var * (M / (M > 1 ? M - 1 : 1)));
  }
}

// ---------------------------------------------------------------------------------------------------------
// BN2 statistics pass: per 256-point unit, h1 = relu(bn1(conv1 x)) (bf16, K-major swizzled operand image), then
// D[channel][point] = W2[channel][k] * h1[point][k]^T on tcgen05 (M = 128 channels, N = 256 points, K = 64), and
// each thread sums its channel's 128 columns.  Two h1 buffers / two TMEM slots: the MMA of unit u overlaps the
// column sums of unit u-1.
struct Stats2Params {
  const float* x;
  int B, P;
  int tiles_per_cloud, total_units;
  const char* packed;      // W2 image at offset 0
  const float* train_par;  // W1p[64][4]
  double *sum2, *sq2;
};
constexpr uint32_t kS2OffW2 = 0, kS2OffH1 = 16384, kS2OffPar = kS2OffH1 + 2 * 32768, kS2OffBar = kS2OffPar + 1024;
constexpr uint32_t kS2Smem = kS2OffBar + 64 + 1024;

__global__ void __launch_bounds__(256, 1) pn_stats2_kernel(const Stats2Params p) {
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

  for (int i = tid; i < 1024; i += 256)
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
  constexpr uint32_t kIdesc = make_idesc(128, 256);
  float rs = 0.f, rq = 0.f;

  for (int it = 0; it <= NU; ++it) {
    if (it < NU) {
      const int unit = u_begin + it;
      const int cloud = unit / p.tiles_per_cloud;
      const int p_base = (unit - cloud * p.tiles_per_cloud) * kUnitPts;
      const float* xc = p.x + (size_t)cloud * 3 * p.P;
      int pt = p_base + tid;
      pt = pt < p.P ? pt : p.P - 1;
      const float x0 = __ldg(xc + pt), x1 = __ldg(xc + p.P + pt), x2 = __ldg(xc + 2 * p.P + pt);
      uint8_t* dst = sm + kS2OffH1 + (it & 1) * 32768;
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
        *reinterpret_cast<uint4*>(dst + sw128_off(tid, cg * 8)) = o;
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (it < NU && warp == 0) {
      if (elect_one()) {
        const uint64_t a_desc = umma_desc_sw128(base + kS2OffW2);
        const uint64_t b_desc = umma_desc_sw128(base + kS2OffH1 + (it & 1) * 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + 256u * (it & 1), a_desc + 2u * k, b_desc + 2u * k, kIdesc, k > 0);
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
      float fs = pair_sum(s2), fq = pair_sum(q2);
      if (ndup > 0) {  // padded columns repeat the last real point: take them out again
        const int valid = kUnitPts - ndup, lo = 128 * hf;
        const int dups = lo + 128 - (valid > lo ? valid : lo);
        if (dups > 0) {
          uint32_t yl;
          tmem_ld1(trow + 255u, yl);
          tmem_ld_wait();
          const float y = __uint_as_float(yl);
          fs -= (float)dups * y;
          fq -= (float)dups * y * y;
        }
      }
      rs += fs;
      rq += fq;
    }
  }
  if (NU > 0) {
    atomicAdd(p.sum2 + q * 32 + lane, (double)rs);
    atomicAdd(p.sq2 + q * 32 + lane, (double)rq);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

}  // namespace pn
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_pointnet_train_ctx_bytes(int64_t B, int64_t P, int64_t F, size_t* bytes) {
  if (!bytes || B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_train_ctx_bytes: bad argument");
  if (!pn::pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  if (B * ((P + pn::kUnitPts - 1) / pn::kUnitPts) >= (1ll << 29) || B * F >= (1ll << 31))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_train_ctx_bytes: problem too large");
  *bytes = pn::TrainCtx((int)B, (int)P, (int)F).total;
  return CRDPN_OK;
}

// Phases (a rank-synchronised run all-reduces the named accumulators between them; see crdpn_pointnet_sync_blocks):
//   0: zero, x moments, operand images            -> sum over ranks: xmom
//   1: BN1 fold, layer-2 statistics pass          -> sum over ranks: sum2 | sq2
//   2: BN2 fold, fused forward                    -> sum over ranks: sum3 | sq3
//   3: BN3 statistics -> output, running stats
extern "C" int crdpn_pointnet_forward_train_phased(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
    const float* conv3_w, const float* conv3_b,
    const float* bn1_w, const float* bn1_b, float* bn1_mean, float* bn1_var, int64_t* bn1_nbt,
    const float* bn2_w, const float* bn2_b, float* bn2_mean, float* bn2_var, int64_t* bn2_nbt,
    const float* bn3_w, const float* bn3_b, float* bn3_mean, float* bn3_var, int64_t* bn3_nbt,
    float bn_eps, float bn_momentum, float* out, void* ctx, size_t ctx_bytes, int variant,
    int phase_begin, int phase_end, int64_t total_points, void* stream) {
  if (!x || !conv1_w || !conv1_b || !conv2_w || !conv2_b || !conv3_w || !conv3_b || !bn1_w || !bn1_b || !bn1_mean ||
      !bn1_var || !bn1_nbt || !bn2_w || !bn2_b || !bn2_mean || !bn2_var || !bn2_nbt || !bn3_w || !bn3_b || !bn3_mean ||
      !bn3_var || !bn3_nbt || !out || !ctx)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_forward_train: null pointer");
  if (B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_forward_train: bad size");
  if (!pn::pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  if (B * ((P + pn::kUnitPts - 1) / pn::kUnitPts) >= (1ll << 29) || B * F >= (1ll << 31))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_forward_train: problem too large");
  if ((uintptr_t)ctx & 1023) return fail(CRDPN_E_ALIGN, "crdpn_pointnet_forward_train: ctx must be 1024-byte aligned");
  const pn::TrainCtx L((int)B, (int)P, (int)F);
  if (ctx_bytes < L.total) return fail(CRDPN_E_WORKSPACE, "crdpn_pointnet_forward_train: ctx too small");
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  if (di.max_smem_optin < (int)pn::kSmemAlloc) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_forward_train: not enough shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  char* c = (char*)ctx;
  // variant bit 4 (16): the bf16 recipe of round 1 (one MMA per product, features within 1e-2 but ~1% of the arg-max / ReLU
  // routing decisions differ from an fp32 run); default: the fp32-accurate split recipe of pointnet_train_split.cu
  const bool bf16_recipe = (variant & 16) != 0;
  if (phase_begin < 0 || phase_end > 4 || phase_begin >= phase_end || total_points < B * P)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_forward_train: bad phase range / total_points");
  auto on = [&](int ph) { return phase_begin <= ph && ph < phase_end; };
  const double M = (double)total_points;  // points of ALL ranks: the batch the statistics are taken over
  double* xmom = (double*)(c + L.xmom);
  double* sum2 = (double*)(c + L.sum2);
  double* sq2 = (double*)(c + L.sq2);
  double* sum3 = (double*)(c + L.sum3);
  double* sq3 = (double*)(c + L.sq3);
  float* stats = (float*)(c + L.stats);
  float* train_par = (float*)(c + L.train_par);

  if (on(0)) {
    CRDPN_CUDA(cudaMemsetAsync(c + L.zero_begin, 0, L.zero_end - L.zero_begin, st));
    pn::pn_xmoments_kernel<<<di.sms * 2, 256, 0, st>>>(x, (int)B, (int)P, xmom);
    CRDPN_LAUNCH_CHECK("pn_xmoments_kernel");
    if (bf16_recipe) {
      pn::pn_pack_train_kernel<<<di.sms, 256, 0, st>>>(conv2_w, conv3_w, bn3_w, (int)F, c + L.packed);
      CRDPN_LAUNCH_CHECK("pn_pack_train_kernel");
    } else {
      rc = pn::split_pack(conv2_w, conv3_w, bn3_w, (int)F, c + L.packed, di.sms, st);
      if (rc) return rc;
    }
  }
  pn::Fold1Params f1{conv1_w, conv1_b, bn1_w, bn1_b, bn1_mean, bn1_var, (long long*)bn1_nbt, (long long*)bn2_nbt,
                     (long long*)bn3_nbt, xmom, M, bn_eps, bn_momentum, stats, train_par, (double*)(c + L.xstat)};
  if (on(1)) {
    pn::pn_fold1_kernel<<<1, 64, 0, st>>>(f1);
    CRDPN_LAUNCH_CHECK("pn_fold1_kernel");
  }

  const int tiles_per_cloud = (int)((P + pn::kUnitPts - 1) / pn::kUnitPts);
  const int total_units = (int)B * tiles_per_cloud;
  const int grid = total_units < di.sms ? total_units : di.sms;
  if (on(1) && !bf16_recipe) {
    rc = pn::split_stats2(x, (int)B, (int)P, c + L.packed, train_par, sum2, sq2, di.sms, st);
    if (rc) return rc;
  } else if (on(1)) {
    static bool attr_set[64] = {false};
    if (!attr_set[device]) {
      CRDPN_CUDA(cudaFuncSetAttribute(pn::pn_stats2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pn::kS2Smem));
      attr_set[device] = true;
    }
    pn::Stats2Params sp{x, (int)B, (int)P, tiles_per_cloud, total_units, c + L.packed, train_par, sum2, sq2};
    pn::pn_stats2_kernel<<<grid, 256, pn::kS2Smem, st>>>(sp);
    CRDPN_LAUNCH_CHECK("pn_stats2_kernel");
  }
  pn::Fold2Params f2{conv2_b, bn2_w, bn2_b, bn2_mean, bn2_var, sum2, sq2, M, bn_eps, bn_momentum, stats, train_par};
  if (on(2)) {
    pn::pn_fold2_kernel<<<1, 128, 0, st>>>(f2);
    CRDPN_LAUNCH_CHECK("pn_fold2_kernel");
  }

  pn::FwdParams fp;
  fp.x = x; fp.B = (int)B; fp.P = (int)P; fp.F = (int)F;
  fp.packed = c + L.packed;
  fp.enc = nullptr;
  fp.tiles_per_cloud = tiles_per_cloud;
  fp.total_units = total_units;
  fp.dbg = nullptr;
  fp.flags = variant & ~(4 | 8 | 16);
  fp.train_par = train_par;
  fp.h2img = c + L.h2img;
  fp.enc64 = (unsigned long long*)(c + L.enc64);
  fp.sum3 = sum3;
  fp.sq3 = sq3;
  if (on(2) && !bf16_recipe) {
    rc = pn::split_forward(x, (int)B, (int)P, (int)F, c + L.packed, train_par, c + L.h2img, (unsigned long long*)(c + L.enc64),
                           sum3, sq3, di.sms, st);
    if (rc) return rc;
  } else if (on(2)) {
    rc = pn::launch_fwd(fp, grid, st);
    if (rc) return rc;
  }

  pn::FinalizeParams fz{(const unsigned long long*)(c + L.enc64), sum3, sq3, conv3_b, bn3_w, bn3_b, bn3_mean, bn3_var,
                        M, bn_eps, bn_momentum, (int)B, (int)F, stats, (int*)(c + L.argmax), (float*)(c + L.yhat3), out};
  const int n = (int)(B * F);
  if (on(3)) {
    pn::pn_train_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(fz);
    CRDPN_LAUNCH_CHECK("pn_train_finalize_kernel");
  }
  return CRDPN_OK;
}

extern "C" int crdpn_pointnet_forward_train(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
    const float* conv3_w, const float* conv3_b,
    const float* bn1_w, const float* bn1_b, float* bn1_mean, float* bn1_var, int64_t* bn1_nbt,
    const float* bn2_w, const float* bn2_b, float* bn2_mean, float* bn2_var, int64_t* bn2_nbt,
    const float* bn3_w, const float* bn3_b, float* bn3_mean, float* bn3_var, int64_t* bn3_nbt,
    float bn_eps, float bn_momentum, float* out, void* ctx, size_t ctx_bytes, int variant, void* stream) {
  return crdpn_pointnet_forward_train_phased(x, B, P, F, conv1_w, conv1_b, conv2_w, conv2_b, conv3_w, conv3_b, bn1_w, bn1_b,
                                             bn1_mean, bn1_var, bn1_nbt, bn2_w, bn2_b, bn2_mean, bn2_var, bn2_nbt, bn3_w,
                                             bn3_b, bn3_mean, bn3_var, bn3_nbt, bn_eps, bn_momentum, out, ctx, ctx_bytes,
                                             variant, 0, 4, B * P, stream);
}

// Accumulators a rank-synchronised (SyncBN-style) run must SUM over ranks after forward phase 0 / 1 / 2 (sync points
// 0..2) and backward phase 0 / 1 / 2 (sync points 3..5).  buffer: 0 = train ctx, 1 = backward workspace, 2 = d_bn3_w,
// 3 = d_bn3_b (the caller's gradient outputs, F floats each).
extern "C" int crdpn_pointnet_sync_blocks(int64_t B, int64_t P, int64_t F, int sync_point, int* n_blocks, int* buffer,
                                          size_t* byte_offset, int64_t* count, int* is_f64) {
  if (!n_blocks || !buffer || !byte_offset || !count || !is_f64 || B <= 0 || P <= 0)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_sync_blocks: bad argument");
  if (!pn::pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  const pn::TrainCtx L((int)B, (int)P, (int)F);
  auto put = [&](int i, int buf, size_t off, int64_t n, int f64) { buffer[i] = buf; byte_offset[i] = off; count[i] = n; is_f64[i] = f64; };
  switch (sync_point) {
    case 0: put(0, 0, L.xmom, 16, 1); *n_blocks = 1; break;
    case 1: put(0, 0, L.sum2, 256, 1); *n_blocks = 1; break;                 // sum2 | sq2 are adjacent
    case 2: put(0, 0, L.sum3, 2 * F, 1); *n_blocks = 1; break;               // sum3 | sq3 are adjacent
    case 3: case 4: case 5: return pn::backward_sync_blocks((int)F, sync_point, n_blocks, buffer, byte_offset, count, is_f64);
    default: return fail(CRDPN_E_BADARG, "crdpn_pointnet_sync_blocks: sync_point must be 0..5");
  }
  return CRDPN_OK;
}
