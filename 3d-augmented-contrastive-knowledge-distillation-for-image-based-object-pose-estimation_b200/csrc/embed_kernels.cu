// embed_kernels.cu -- the CRD embed heads (published `Embed`: flatten -> Linear(dim_in, D) -> x / ||x||_2),
// forward and backward, as a handful of small kernels instead of the ~25 library / elementwise launches the eager
// module costs per step (B = 46 rows: the work is launch-bound, not FLOP-bound).
//   forward : pre = x W^T + b (warp per output, float4 loads);  v = pre / ||pre||_2, 1/norm kept for backward
//   backward: d_pre = (g - v (g . v)) / norm;  dW = d_pre^T x;  db = sum_b d_pre;  dx = d_pre W (optional)
#include "common.cuh"

namespace crdpn {
namespace embed {

// grid (ceil(D/8), B), 256 threads: warp w of block (bx, b) computes output d = 8 bx + w of row b
__device__ __forceinline__ void embed_linear_body(const float* __restrict__ x, const float* __restrict__ W,
                                                  const float* __restrict__ bias, int B, int dim_in, int D,
                                                  float* __restrict__ pre, int bx, int by) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = bx * 8 + warp;
  if (d >= D) return;
  const float* wr = W + (size_t)d * dim_in;
  const int b0 = by, b1 = min(b0 + 1, B);
  const bool vec = (dim_in & 3) == 0;
  for (int b = b0; b < b1; ++b) {
    const float* xr = x + (size_t)b * dim_in;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (vec) {
      for (int i = lane * 4; i < dim_in; i += 128) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + i));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
        a0 = fmaf(xv.x, wv.x, a0); a1 = fmaf(xv.y, wv.y, a1); a2 = fmaf(xv.z, wv.z, a2); a3 = fmaf(xv.w, wv.w, a3);
      }
    } else {
      for (int i = lane; i < dim_in; i += 32) a0 = fmaf(__ldg(xr + i), __ldg(wr + i), a0);
    }
    float s = (a0 + a1) + (a2 + a3);
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) pre[(size_t)b * D + d] = s + bias[d];
  }
}
__global__ void __launch_bounds__(256) embed_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ bias, int B, int dim_in, int D,
                                                           float* __restrict__ pre) {
  embed_linear_body(x, W, bias, B, dim_in, D, pre, blockIdx.x, blockIdx.y);
}
// both heads in one launch: grid (ceil(D/8), B, 2), blockIdx.z = head
struct Head { const float* x; const float* W; const float* b; int dim_in; float* pre; float* v; float* inv; };
__global__ void __launch_bounds__(256) embed_linear2_kernel(const Head h0, const Head h1, int B, int D) {
  const Head& h = blockIdx.z == 0 ? h0 : h1;
  embed_linear_body(h.x, h.W, h.b, B, h.dim_in, D, h.pre, blockIdx.x, blockIdx.y);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}

// grid B, 256 threads: v = pre / ||pre||
__device__ __forceinline__ void embed_normalize_body(const float* __restrict__ pre, int D, float* __restrict__ v,
                                                     float* __restrict__ inv_norm, int row, float* red) {
  const float* pr = pre + (size_t)row * D;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) ss = fmaf(pr[d], pr[d], ss);
  ss = block_sum(ss, red);
  const float norm = sqrtf(ss);
  for (int d = threadIdx.x; d < D; d += blockDim.x) v[(size_t)row * D + d] = pr[d] / norm;
  if (threadIdx.x == 0) inv_norm[row] = 1.0f / norm;
}
__global__ void __launch_bounds__(256) embed_normalize_kernel(const float* __restrict__ pre, int D, float* __restrict__ v,
                                                              float* __restrict__ inv_norm) {
  __shared__ float red[8];
  embed_normalize_body(pre, D, v, inv_norm, blockIdx.x, red);
}
// grid (B, 2)
__global__ void __launch_bounds__(256) embed_normalize2_kernel(const Head h0, const Head h1, int D) {
  __shared__ float red[8];
  const Head& h = blockIdx.y == 0 ? h0 : h1;
  embed_normalize_body(h.pre, D, h.v, h.inv, blockIdx.x, red);
}

// grid B, 256 threads: d_pre = scale * (g - v (g . v)) / norm   (scale: optional device scalar, the upstream gradient)
__device__ __forceinline__ void embed_bwd_prep_body(const float* __restrict__ g, const float* __restrict__ v,
                                                    const float* __restrict__ inv_norm, const float* __restrict__ scale,
                                                    int D, float* __restrict__ d_pre, int row, float* red) {
  const size_t o = (size_t)row * D;
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dot = fmaf(g[o + d], v[o + d], dot);
  dot = block_sum(dot, red);
  const float inv = inv_norm[row] * (scale ? *scale : 1.0f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) d_pre[o + d] = (g[o + d] - v[o + d] * dot) * inv;
}
__global__ void __launch_bounds__(256) embed_bwd_prep_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                                             const float* __restrict__ inv_norm, const float* __restrict__ scale,
                                                             int D, float* __restrict__ d_pre) {
  __shared__ float red[8];
  embed_bwd_prep_body(g, v, inv_norm, scale, D, d_pre, blockIdx.x, red);
}
struct BwdHead { const float* x; const float* W; const float* v; const float* inv; const float* g; int dim_in;
                 float* dW; float* db; float* dx; float* d_pre; };
// grid (B, 2)
__global__ void __launch_bounds__(256) embed_bwd_prep2_kernel(const BwdHead h0, const BwdHead h1, const float* __restrict__ scale, int D) {
  __shared__ float red[8];
  const BwdHead& h = blockIdx.y == 0 ? h0 : h1;
  embed_bwd_prep_body(h.g, h.v, h.inv, scale, D, h.d_pre, blockIdx.x, red);
}

// grid (ceil(D/4), ceil(dim_in/256)), 256 threads: dW[d][i] = sum_b d_pre[b][d] x[b][i];  db[d] = sum_b d_pre[b][d].
// Each thread keeps FOUR outputs (d0..d0+3, same column i): one coalesced x load feeds four FMAs, so the L2 traffic on x
// (the operand every block re-reads) is a quarter of the one-output version's.
constexpr int kEmbedR = 4;
__device__ __forceinline__ void embed_bwd_wgrad_body(const float* __restrict__ d_pre, const float* __restrict__ x,
                                                     int B, int dim_in, int D, float* __restrict__ dW, float* __restrict__ db,
                                                     int bx, int by) {
  const int d0 = bx * kEmbedR, i = by * 256 + threadIdx.x;
  if (i < dim_in) {
    float acc[kEmbedR] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      const float xv = __ldg(x + (size_t)b * dim_in + i);
#pragma unroll
      for (int q = 0; q < kEmbedR; ++q)
        if (d0 + q < D) acc[q] = fmaf(__ldg(d_pre + (size_t)b * D + d0 + q), xv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < kEmbedR; ++q)
      if (d0 + q < D) dW[(size_t)(d0 + q) * dim_in + i] = acc[q];
  }
  if (by == 0 && threadIdx.x < kEmbedR && d0 + (int)threadIdx.x < D) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += d_pre[(size_t)b * D + d0 + threadIdx.x];
    db[d0 + threadIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) embed_bwd_wgrad_kernel(const float* __restrict__ d_pre, const float* __restrict__ x,
                                                              int B, int dim_in, int D, float* __restrict__ dW, float* __restrict__ db) {
  embed_bwd_wgrad_body(d_pre, x, B, dim_in, D, dW, db, blockIdx.x, blockIdx.y);
}

// grid (ceil(B/4), ceil(dim_in/256)), 256 threads: dx[b][i] = sum_d d_pre[b][d] W[d][i], four rows b per thread (one
// coalesced W load feeds four FMAs)
__device__ __forceinline__ void embed_bwd_dgrad_body(const float* __restrict__ d_pre, const float* __restrict__ W,
                                                     int B, int dim_in, int D, float* __restrict__ dx, int bx, int by) {
  const int b0 = bx * kEmbedR, i = by * 256 + threadIdx.x;
  if (i >= dim_in) return;
  float acc[kEmbedR] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
    const float wv = __ldg(W + (size_t)d * dim_in + i);
#pragma unroll
    for (int q = 0; q < kEmbedR; ++q)
      if (b0 + q < B) acc[q] = fmaf(__ldg(d_pre + (size_t)(b0 + q) * D + d), wv, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < kEmbedR; ++q)
    if (b0 + q < B) dx[(size_t)(b0 + q) * dim_in + i] = acc[q];
}
__global__ void __launch_bounds__(256) embed_bwd_dgrad_kernel(const float* __restrict__ d_pre, const float* __restrict__ W,
                                                              int B, int dim_in, int D, float* __restrict__ dx) {
  embed_bwd_dgrad_body(d_pre, W, B, dim_in, D, dx, blockIdx.x, blockIdx.y);
}
// wgrad and dgrad of BOTH heads in one launch (they only depend on d_pre): 1-D grid cut into four block ranges
// [wgrad head 0 | wgrad head 1 | dgrad head 0 | dgrad head 1]; n_* = blocks of each range (0 when that gradient is off)
__global__ void __launch_bounds__(256) embed_bwd_grads2_kernel(const BwdHead h0, const BwdHead h1, int B, int D,
                                                               int tiles0, int tiles1, int n_w0, int n_w1, int n_d0) {
  int blk = blockIdx.x;
  if (blk < n_w0) { embed_bwd_wgrad_body(h0.d_pre, h0.x, B, h0.dim_in, D, h0.dW, h0.db, blk / tiles0, blk % tiles0); return; }
  blk -= n_w0;
  if (blk < n_w1) { embed_bwd_wgrad_body(h1.d_pre, h1.x, B, h1.dim_in, D, h1.dW, h1.db, blk / tiles1, blk % tiles1); return; }
  blk -= n_w1;
  if (blk < n_d0) { embed_bwd_dgrad_body(h0.d_pre, h0.W, B, h0.dim_in, D, h0.dx, blk / tiles0, blk % tiles0); return; }
  blk -= n_d0;
  embed_bwd_dgrad_body(h1.d_pre, h1.W, B, h1.dim_in, D, h1.dx, blk / tiles1, blk % tiles1);
}

}  // namespace embed
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_embed_forward(const float* x, const float* W, const float* b, int64_t B, int64_t dim_in, int64_t D,
                                   float* pre, float* v, float* inv_norm, void* stream) {
  if (!x || !W || !b || !pre || !v || !inv_norm) return fail(CRDPN_E_BADARG, "crdpn_embed_forward: null pointer");
  if (B <= 0 || dim_in <= 0 || D <= 0 || B > 65535 || dim_in >= (1ll << 31) || D >= (1ll << 24))
    return fail(CRDPN_E_BADARG, "crdpn_embed_forward: bad size");
  if ((dim_in & 3) == 0 && (((uintptr_t)x | (uintptr_t)W) & 15)) return fail(CRDPN_E_ALIGN, "crdpn_embed_forward: x / W must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((D + 7) / 8), (unsigned)B);
  embed::embed_linear_kernel<<<grid, 256, 0, st>>>(x, W, b, (int)B, (int)dim_in, (int)D, pre);
  CRDPN_LAUNCH_CHECK("embed_linear_kernel");
  embed::embed_normalize_kernel<<<(unsigned)B, 256, 0, st>>>(pre, (int)D, v, inv_norm);
  CRDPN_LAUNCH_CHECK("embed_normalize_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_embed_backward(const float* x, const float* W, const float* v, const float* inv_norm,
                                    const float* grad_v, const float* scale, int64_t B, int64_t dim_in, int64_t D,
                                    float* dW, float* db, float* dx, float* d_pre, void* stream) {
  if (!x || !W || !v || !inv_norm || !grad_v || !dW || !db || !d_pre) return fail(CRDPN_E_BADARG, "crdpn_embed_backward: null pointer");
  if (B <= 0 || dim_in <= 0 || D <= 0 || B > 65535 || D > 65535 * 1 || dim_in >= (1ll << 31))
    return fail(CRDPN_E_BADARG, "crdpn_embed_backward: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  embed::embed_bwd_prep_kernel<<<(unsigned)B, 256, 0, st>>>(grad_v, v, inv_norm, scale, (int)D, d_pre);
  CRDPN_LAUNCH_CHECK("embed_bwd_prep_kernel");
  const unsigned tiles = (unsigned)((dim_in + 255) / 256);
  embed::embed_bwd_wgrad_kernel<<<dim3((unsigned)((D + embed::kEmbedR - 1) / embed::kEmbedR), tiles), 256, 0, st>>>(d_pre, x, (int)B, (int)dim_in, (int)D, dW, db);
  CRDPN_LAUNCH_CHECK("embed_bwd_wgrad_kernel");
  if (dx) {
    embed::embed_bwd_dgrad_kernel<<<dim3((unsigned)((B + embed::kEmbedR - 1) / embed::kEmbedR), tiles), 256, 0, st>>>(d_pre, W, (int)B, (int)dim_in, (int)D, dx);
    CRDPN_LAUNCH_CHECK("embed_bwd_dgrad_kernel");
  }
  return CRDPN_OK;
}

// Both embed heads of CRDLoss at once: 2 launches forward, 2 backward (instead of 4 and 5-6).  Same arithmetic, same
// bits as the per-head calls.
namespace crdpn {
int embed_forward2(const float* xs, const float* Ws, const float* bs, int64_t s_dim, float* pre_s, float* v1, float* inv1,
                   const float* xt, const float* Wt, const float* bt, int64_t t_dim, float* pre_t, float* v2, float* inv2,
                   int64_t B, int64_t D, void* stream) {
  if (!xs || !Ws || !bs || !pre_s || !v1 || !inv1 || !xt || !Wt || !bt || !pre_t || !v2 || !inv2)
    return fail(CRDPN_E_BADARG, "embed_forward2: null pointer");
  if (B <= 0 || s_dim <= 0 || t_dim <= 0 || D <= 0 || B > 65535 || s_dim >= (1ll << 31) || t_dim >= (1ll << 31) || D >= (1ll << 24))
    return fail(CRDPN_E_BADARG, "embed_forward2: bad size");
  if (((s_dim & 3) == 0 && (((uintptr_t)xs | (uintptr_t)Ws) & 15)) || ((t_dim & 3) == 0 && (((uintptr_t)xt | (uintptr_t)Wt) & 15)))
    return fail(CRDPN_E_ALIGN, "embed_forward2: x / W must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const embed::Head h0{xs, Ws, bs, (int)s_dim, pre_s, v1, inv1}, h1{xt, Wt, bt, (int)t_dim, pre_t, v2, inv2};
  embed::embed_linear2_kernel<<<dim3((unsigned)((D + 7) / 8), (unsigned)B, 2), 256, 0, st>>>(h0, h1, (int)B, (int)D);
  CRDPN_LAUNCH_CHECK("embed_linear2_kernel");
  embed::embed_normalize2_kernel<<<dim3((unsigned)B, 2), 256, 0, st>>>(h0, h1, (int)D);
  CRDPN_LAUNCH_CHECK("embed_normalize2_kernel");
  return CRDPN_OK;
}

int embed_backward2(const float* xs, int64_t s_dim, const float* Ws, const float* v1, const float* inv1, const float* g1,
                    const float* xt, int64_t t_dim, const float* Wt, const float* v2, const float* inv2, const float* g2,
                    const float* scale, int64_t B, int64_t D, float* dWs, float* dbs, float* dxs, float* dWt, float* dbt,
                    float* dxt, float* d_pre, void* stream) {
  if (!xs || !Ws || !v1 || !inv1 || !g1 || !xt || !Wt || !v2 || !inv2 || !g2 || !dWs || !dbs || !dWt || !dbt || !d_pre)
    return fail(CRDPN_E_BADARG, "embed_backward2: null pointer");
  if (B <= 0 || s_dim <= 0 || t_dim <= 0 || D <= 0 || B > 65535 || D > 65535 || s_dim >= (1ll << 31) || t_dim >= (1ll << 31))
    return fail(CRDPN_E_BADARG, "embed_backward2: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  const embed::BwdHead h0{xs, Ws, v1, inv1, g1, (int)s_dim, dWs, dbs, dxs, d_pre};
  const embed::BwdHead h1{xt, Wt, v2, inv2, g2, (int)t_dim, dWt, dbt, dxt, d_pre + B * D};
  embed::embed_bwd_prep2_kernel<<<dim3((unsigned)B, 2), 256, 0, st>>>(h0, h1, scale, (int)D);
  CRDPN_LAUNCH_CHECK("embed_bwd_prep2_kernel");
  const long long t0 = (s_dim + 255) / 256, t1 = (t_dim + 255) / 256;
  const long long dg = (D + embed::kEmbedR - 1) / embed::kEmbedR, bg = (B + embed::kEmbedR - 1) / embed::kEmbedR;
  const long long n_w0 = dg * t0, n_w1 = dg * t1, n_d0 = dxs ? bg * t0 : 0, n_d1 = dxt ? bg * t1 : 0;
  const long long total = n_w0 + n_w1 + n_d0 + n_d1;
  if (total >= (1ll << 31)) return fail(CRDPN_E_UNSUPPORTED, "embed_backward2: grid too large");
  embed::embed_bwd_grads2_kernel<<<(unsigned)total, 256, 0, st>>>(h0, h1, (int)B, (int)D, (int)t0, (int)t1, (int)n_w0, (int)n_w1,
                                                                 (int)n_d0);
  CRDPN_LAUNCH_CHECK("embed_bwd_grads2_kernel");
  return CRDPN_OK;
}
}  // namespace crdpn
