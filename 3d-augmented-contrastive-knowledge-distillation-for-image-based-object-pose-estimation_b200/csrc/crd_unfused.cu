// crd_unfused.cu -- backward of the UNFUSED published surface ContrastMemory.forward -> (out_v1, out_v2).
//
// CRDLoss (the call the KD loop makes) never needs this: its fused kernel applies the closed-form NCE gradient in the
// scoring pass.  Users who call ContrastMemory themselves and put their own criterion on out_v1 / out_v2 get gradients
// through this kernel instead:
//     out_v1[b,k] = exp(<bank2[idx[b,k]], v1[b]> / T) / Z1   =>   grad_v1[b] = sum_k go1[b,k] * out_v1[b,k] / T * bank2[idx[b,k]]
//     out_v2[b,k] = exp(<bank1[idx[b,k]], v2[b]> / T) / Z2   =>   grad_v2[b] = sum_k go2[b,k] * out_v2[b,k] / T * bank1[idx[b,k]]
// The published code detaches a COPY of the gathered rows, so its backward sees the banks as they were BEFORE the
// momentum update of the same forward call; the rows `y` that the update overwrote are therefore passed in as saved
// copies (old1 / old2 [B, D]) and substituted wherever a contrast index hits one of them (shared-memory hash of y).
// Same streaming pattern as the scoring pass (one 128-bit load per lane per row), deterministic two-stage reduction.
#include "common.cuh"

namespace crdpn {
namespace unf {

constexpr int kThreads = 256;
constexpr int kChunk = 1024;     // contrast entries per CTA
constexpr int kMaxVec = 4;       // float4 accumulators per lane and bank: D <= 512
constexpr int kHashSlots = 2048; // >= 2 * B

struct Params {
  const char* bank1;
  const char* bank2;
  long long row_stride_bytes;
  int bf16;
  const float* old1;  // [B, D] pre-update rows of y (fp32)
  const float* old2;
  const long long* y;
  const long long* idx;
  const float* go1; const float* go2;  // [B, K1]
  const float* o1; const float* o2;    // [B, K1]
  int B, K1, D, S;
  long long row_begin, row_end;
  float inv_T;
  float* partial;  // [B, S, 2, D]
};

__device__ __forceinline__ float4 load_row4(const char* bank, long long row, long long stride_bytes, int e, int bf16) {
  if (!bf16) return __ldg(reinterpret_cast<const float4*>(bank + row * stride_bytes) + (e >> 2));
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(bank + row * stride_bytes) + (e >> 2));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

__global__ void __launch_bounds__(kThreads) crd_out_backward_kernel(const Params p) {
  __shared__ long long h_key[kHashSlots];
  __shared__ int h_val[kHashSlots];
  __shared__ float4 red[kThreads / 32][2][32];
  const int b = blockIdx.y, s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) h_key[i] = -1;
  __syncthreads();
  if (threadIdx.x == 0) {  // B <= 1024 insertions; last occurrence wins (duplicates hold identical saved rows anyway)
    for (int j = 0; j < p.B; ++j) {
      const long long key = p.y[j];
      unsigned slot = (unsigned)((unsigned long long)key * 0x9E3779B97F4A7C15ull >> 53) & (kHashSlots - 1);
      while (h_key[slot] != -1 && h_key[slot] != key) slot = (slot + 1) & (kHashSlots - 1);
      h_key[slot] = key;
      h_val[slot] = j;
    }
  }
  __syncthreads();
  const int nvec = (p.D + 127) / 128;
  float4 a1[kMaxVec], a2[kMaxVec];
#pragma unroll
  for (int v = 0; v < kMaxVec; ++v) a1[v] = a2[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int k0 = s * kChunk, k1 = min(k0 + kChunk, p.K1);
  for (int k = k0 + warp; k < k1; k += kThreads / 32) {
    const size_t pos = (size_t)b * p.K1 + k;
    const long long row = p.idx[pos];
    if (row < p.row_begin || row >= p.row_end) continue;
    const float c1 = p.go1[pos] * p.o1[pos] * p.inv_T, c2 = p.go2[pos] * p.o2[pos] * p.inv_T;
    unsigned slot = (unsigned)((unsigned long long)row * 0x9E3779B97F4A7C15ull >> 53) & (kHashSlots - 1);
    int saved = -1;
    while (h_key[slot] != -1) {
      if (h_key[slot] == row) { saved = h_val[slot]; break; }
      slot = (slot + 1) & (kHashSlots - 1);
    }
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      const int e = v * 128 + lane * 4;
      if (v < nvec && e < p.D) {
        float4 r1, r2;
        if (saved >= 0) {
          r1 = *reinterpret_cast<const float4*>(p.old1 + (size_t)saved * p.D + e);
          r2 = *reinterpret_cast<const float4*>(p.old2 + (size_t)saved * p.D + e);
        } else {
          r1 = load_row4(p.bank1, row - p.row_begin, p.row_stride_bytes, e, p.bf16);
          r2 = load_row4(p.bank2, row - p.row_begin, p.row_stride_bytes, e, p.bf16);
        }
        a1[v].x = fmaf(c1, r2.x, a1[v].x); a1[v].y = fmaf(c1, r2.y, a1[v].y);
        a1[v].z = fmaf(c1, r2.z, a1[v].z); a1[v].w = fmaf(c1, r2.w, a1[v].w);
        a2[v].x = fmaf(c2, r1.x, a2[v].x); a2[v].y = fmaf(c2, r1.y, a2[v].y);
        a2[v].z = fmaf(c2, r1.z, a2[v].z); a2[v].w = fmaf(c2, r1.w, a2[v].w);
      }
    }
  }
  float* out = p.partial + ((size_t)b * p.S + s) * 2 * p.D;
  for (int v = 0; v < nvec; ++v) {
    __syncthreads();
    red[warp][0][lane] = a1[v];
    red[warp][1][lane] = a2[v];
    __syncthreads();
    if (warp < 2) {  // warp 0 folds grad_v1's slice, warp 1 grad_v2's, over the 8 warps in fixed order
      float4 acc = red[0][warp][lane];
      for (int w = 1; w < kThreads / 32; ++w) {
        const float4 t = red[w][warp][lane];
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      const int e = v * 128 + lane * 4;
      if (e < p.D) *reinterpret_cast<float4*>(out + (size_t)warp * p.D + e) = acc;
    }
  }
}

// grid B, threads over 2*D: fixed-order sum of the S chunk partials
__global__ void __launch_bounds__(256) crd_out_backward_reduce_kernel(const float* __restrict__ partial, int S, int D,
                                                                      float* __restrict__ g1, float* __restrict__ g2) {
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < 2 * D; e += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += partial[((size_t)b * S + s) * 2 * D + e];
    if (e < D) g1[(size_t)b * D + e] = acc;
    else g2[(size_t)b * D + (e - D)] = acc;
  }
}

}  // namespace unf
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_crd_out_backward_workspace_bytes(int64_t B, int64_t K1, int64_t D, size_t* bytes) {
  if (!bytes || B <= 0 || K1 <= 0 || D <= 0) return fail(CRDPN_E_BADARG, "crdpn_crd_out_backward_workspace_bytes: bad argument");
  const int64_t S = (K1 + unf::kChunk - 1) / unf::kChunk;
  *bytes = (size_t)B * S * 2 * D * sizeof(float);
  return CRDPN_OK;
}

extern "C" int crdpn_crd_out_backward(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                                      const float* old_rows1, const float* old_rows2, const int64_t* y,
                                      const int64_t* contrast_idx, const float* grad_out_v1, const float* grad_out_v2,
                                      const float* out_v1, const float* out_v2, int64_t B, int64_t K1, int64_t D,
                                      int64_t row_begin, int64_t row_end, float T, float* grad_v1, float* grad_v2,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!old_rows1 || !old_rows2 || !y || !contrast_idx || !grad_out_v1 || !grad_out_v2 || !out_v1 || !out_v2 || !grad_v1 ||
      !grad_v2 || !workspace)
    return fail(CRDPN_E_BADARG, "crdpn_crd_out_backward: null pointer");
  if (B <= 0 || K1 <= 0 || D <= 0 || row_end < row_begin || !(T > 0.f)) return fail(CRDPN_E_BADARG, "crdpn_crd_out_backward: bad size");
  if (row_end > row_begin && (!bank1 || !bank2)) return fail(CRDPN_E_BADARG, "crdpn_crd_out_backward: null bank");
  if (D % 4 != 0 || D > 128 * unf::kMaxVec || 2 * B > unf::kHashSlots || B > 65535)
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_out_backward: feat_dim must be a multiple of 4 and <= 512, batch <= 1024");
  if (bank_dtype != CRDPN_F32 && bank_dtype != CRDPN_BF16) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_out_backward: bank dtype");
  const size_t esz = bank_dtype == CRDPN_BF16 ? 2 : 4;
  if (((uintptr_t)bank1 | (uintptr_t)bank2 | (uintptr_t)old_rows1 | (uintptr_t)old_rows2 | (uintptr_t)workspace) & 15 ||
      ((size_t)row_stride * esz) % 16 != 0)
    return fail(CRDPN_E_ALIGN, "crdpn_crd_out_backward: 16-byte alignment required");
  const int64_t S = (K1 + unf::kChunk - 1) / unf::kChunk;
  if (workspace_bytes < (size_t)B * S * 2 * D * sizeof(float)) return fail(CRDPN_E_WORKSPACE, "crdpn_crd_out_backward: workspace too small");
  unf::Params p;
  p.bank1 = (const char*)bank1; p.bank2 = (const char*)bank2;
  p.row_stride_bytes = (long long)row_stride * (long long)esz;
  p.bf16 = bank_dtype == CRDPN_BF16;
  p.old1 = old_rows1; p.old2 = old_rows2;
  p.y = (const long long*)y; p.idx = (const long long*)contrast_idx;
  p.go1 = grad_out_v1; p.go2 = grad_out_v2; p.o1 = out_v1; p.o2 = out_v2;
  p.B = (int)B; p.K1 = (int)K1; p.D = (int)D; p.S = (int)S;
  p.row_begin = row_begin; p.row_end = row_end;
  p.inv_T = 1.0f / T;
  p.partial = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  unf::crd_out_backward_kernel<<<dim3((unsigned)S, (unsigned)B), unf::kThreads, 0, st>>>(p);
  CRDPN_LAUNCH_CHECK("crd_out_backward_kernel");
  unf::crd_out_backward_reduce_kernel<<<(unsigned)B, 256, 0, st>>>(p.partial, (int)S, (int)D, grad_v1, grad_v2);
  CRDPN_LAUNCH_CHECK("crd_out_backward_reduce_kernel");
  return CRDPN_OK;
}
