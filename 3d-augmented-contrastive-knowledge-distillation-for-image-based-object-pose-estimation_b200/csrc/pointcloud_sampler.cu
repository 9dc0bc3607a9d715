// pointcloud_sampler.cu -- the input producer of the PointNet encoder, for a whole batch in ONE launch.
//
// Reference: read_pointcloud, auxiliary/dataset.py:121-150 (called per sample from DataLoader workers,
// dataset.py:299,607): load the mesh vertices, pick `point_num` of them without replacement, optionally rotate about z,
// transpose to [3, P] float32, shift by the global minimum and divide by the global maximum so that the cloud lies in
// [0, 1].  Here the raw vertices of every CAD model stay resident in HBM (a few hundred models x 10^4 vertices x 24 B),
// and a step's B clouds are produced by one kernel: one CTA per cloud, every value written once unnormalised, the CTA's
// min / max reduced on chip, then the same threads re-read their own values and normalise.  No atomics, no workspace.
//
// The subset is either given ([B, P] indices: parity runs) or generated in the kernel as the first P images of a keyed
// pseudo-random PERMUTATION of [0, V): a 6-round Feistel network on the smallest even bit width covering V with cycle
// walking -- O(1) per point, no memory, distinct by construction.  oracle/pointcloud_oracle.py holds the same function.
#include "common.cuh"

namespace crdpn {
namespace pcs {

constexpr int kThreads = 512;
constexpr int kMaxPerThread = 16;  // register-resident values per thread and coordinate: P <= 8192

struct FeistelKey {
  unsigned k[6];
  int half_bits;       // the permutation acts on 2 * half_bits bits
};

__host__ __device__ inline unsigned feistel_round(unsigned r, unsigned key) {
  unsigned x = r * 0x9E3779B1u + key;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}

// image of i under the keyed permutation of [0, V)
__host__ __device__ inline unsigned long long feistel_perm(unsigned long long i, unsigned long long V, const FeistelKey& fk) {
  const unsigned mask = (fk.half_bits >= 32) ? 0xFFFFFFFFu : ((1u << fk.half_bits) - 1u);
  unsigned long long x = i;
  do {
    unsigned L = (unsigned)(x >> fk.half_bits) & mask, R = (unsigned)x & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const unsigned t = L ^ (feistel_round(R, fk.k[r]) & mask);
      L = R;
      R = t;
    }
    x = ((unsigned long long)L << fk.half_bits) | (unsigned long long)R;
  } while (x >= V);
  return x;
}

__device__ __forceinline__ FeistelKey make_key(unsigned long long seed, unsigned long long stream, unsigned long long V) {
  FeistelKey fk;
  unsigned a[4], b[4];
  philox4x32_10(seed, 2ull * stream, a);
  philox4x32_10(seed, 2ull * stream + 1ull, b);
  fk.k[0] = a[0]; fk.k[1] = a[1]; fk.k[2] = a[2]; fk.k[3] = a[3]; fk.k[4] = b[0]; fk.k[5] = b[1];
  int bits = 2;
  while (bits < 64 && (1ull << bits) < V) bits += 2;
  fk.half_bits = bits / 2;
  return fk;
}

__device__ __forceinline__ float block_min(float v, float* red) {
  for (int off = 16; off >= 1; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s = fminf(s, red[w]);
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s = fmaxf(s, red[w]);
  return s;
}

// grid B, 512 threads
__global__ void __launch_bounds__(kThreads) pointcloud_sample_kernel(
    const double* __restrict__ vertices, const long long* __restrict__ cloud_offsets, const long long* __restrict__ cloud_ids,
    const float* __restrict__ rotation_deg, const long long* __restrict__ subset, unsigned long long seed,
    unsigned long long offset, int P, float* __restrict__ out, long long* __restrict__ subset_out) {
  __shared__ float red[kThreads / 32];
  const int b = blockIdx.x;
  const long long cid = cloud_ids[b];
  const long long v0 = cloud_offsets[cid];
  const unsigned long long V = (unsigned long long)(cloud_offsets[cid + 1] - v0);
  float* ob = out + (size_t)b * 3 * P;
  if (V < (unsigned long long)P) {  // the reference raises (np.random.choice without replacement); mark the cloud invalid
    for (int i = threadIdx.x; i < 3 * P; i += blockDim.x) ob[i] = __int_as_float(0x7fc00000);
    return;
  }
  FeistelKey fk;
  if (subset == nullptr) fk = make_key(seed, offset + (unsigned long long)b, V);
  const double rot = rotation_deg ? (double)rotation_deg[b] : 0.0;
  const bool rotate = rot != 0.0;
  double c = 1.0, s = 0.0;
  if (rotate) {
    const double alpha = rot * 0.017453292519943295;  // math.radians
    c = cos(alpha);
    s = sin(alpha);
  }
  float vx[kMaxPerThread], vy[kMaxPerThread], vz[kMaxPerThread];
  float lo = INFINITY, hi = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxPerThread; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < P) {
      unsigned long long src;
      if (subset != nullptr) {
        const long long sidx = subset[(size_t)b * P + i];
        src = (sidx < 0 || (unsigned long long)sidx >= V) ? 0ull : (unsigned long long)sidx;
      } else {
        src = feistel_perm((unsigned long long)i, V, fk);
      }
      if (subset_out != nullptr) subset_out[(size_t)b * P + i] = (long long)src;
      const double* vp = vertices + 3 * (size_t)(v0 + (long long)src);
      double x = vp[0], y = vp[1];
      const double z = vp[2];
      if (rotate) {  // point_cloud @ rot_matrix.T (dataset.py:138-142), float64 as there
        const double xr = x * c + y * (-s) + z * 0.0;
        const double yr = x * s + y * c + z * 0.0;
        x = xr;
        y = yr;
      }
      vx[j] = (float)x; vy[j] = (float)y; vz[j] = (float)z;
      lo = fminf(lo, fminf(vx[j], fminf(vy[j], vz[j])));
      hi = fmaxf(hi, fmaxf(vx[j], fmaxf(vy[j], vz[j])));
    }
  }
  lo = block_min(lo, red);
  hi = block_max(hi, red);
  const float span = hi - lo;  // == max(cloud - min) in fp32: the subtraction is monotone
#pragma unroll
  for (int j = 0; j < kMaxPerThread; ++j) {
    const int i = threadIdx.x + j * kThreads;
    if (i < P) {
      ob[i] = (vx[j] - lo) / span;
      ob[P + i] = (vy[j] - lo) / span;
      ob[2 * (size_t)P + i] = (vz[j] - lo) / span;
    }
  }
}

}  // namespace pcs
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_pointcloud_sample(const double* vertices, const int64_t* cloud_offsets, const int64_t* cloud_ids,
                                       const float* rotation_deg, const int64_t* subset, uint64_t seed, uint64_t offset,
                                       int64_t B, int64_t P, float* out, int64_t* subset_out, void* stream) {
  if (!vertices || !cloud_offsets || !cloud_ids || !out) return fail(CRDPN_E_BADARG, "crdpn_pointcloud_sample: null pointer");
  if (B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointcloud_sample: bad size");
  if (P > pcs::kThreads * pcs::kMaxPerThread)
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointcloud_sample: point_num above 8192 (the per-thread register tile)");
  pcs::pointcloud_sample_kernel<<<(unsigned)B, pcs::kThreads, 0, (cudaStream_t)stream>>>(
      vertices, (const long long*)cloud_offsets, (const long long*)cloud_ids, rotation_deg, (const long long*)subset, seed, offset,
      (int)P, out, (long long*)subset_out);
  CRDPN_LAUNCH_CHECK("pointcloud_sample_kernel");
  return CRDPN_OK;
}
