// pointnet_train.cuh -- layout of the caller-owned train-mode context buffer shared by the train forward
// (pointnet_train.cu) and the backward (pointnet_backward.cu).  All offsets in bytes from a 1024-aligned base.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>

namespace crdpn {
namespace pn {

// float offsets inside the statistics block: batch mean of each conv output WITHOUT its bias (true sign) and
// 1/sqrt(biased batch variance + eps)
constexpr int kStatMean1 = 0, kStatIstd1 = 64, kStatMean2 = 128, kStatIstd2 = 256, kStatMean3 = 384;
__host__ __device__ inline int kStatIstd3(int F) { return kStatMean3 + F; }
__host__ __device__ inline int kStatFloats(int F) { return kStatMean3 + 2 * F; }

struct TrainCtx {
  // zeroed at the start of every forward: [zero_begin, zero_end)
  size_t zero_begin, xmom, sum2, sq2, sum3, sq3, enc64, zero_end;
  size_t xstat;      // double[12]: mean x (3), Cov x (3x3)
  size_t stats;      // float[kStatFloats(F)]
  size_t train_par;  // float[512]: W1p[64][4] (BN1 folded), sh2[128], sc2[128]
  size_t argmax;     // int32[B*F]: point index of the max per (cloud, channel)
  size_t yhat3;      // float[B*F]: normalised conv3 output at that point, (y3 - mean3) * istd3
  size_t packed;     // operand images of the forward: fp16 hi/lo pieces (split recipe: W2 hi | lo, 32 KB, then per slab
                     // hi k0 | hi k1 | lo k0 | lo k1, F/128 x 64 KB) or bf16 (W2 16 KB | sign(gamma3)*W3 slabs F/128 x 32 KB)
  size_t h2img;      // bf16 h2 tiles: [B * tiles2][32 KB], tiles2 = 2*ceil(P/256) tiles of 128 points per cloud
  size_t total;
  int tiles2;        // 128-point tiles per cloud (even; trailing rows/tiles past P repeat the last point)

  __host__ __device__ TrainCtx(int B, int P, int F) {
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t o = 0;
    zero_begin = o;
    xmom = o; o += 16 * 8;
    sum2 = o; o += 128 * 8;
    sq2 = o; o += 128 * 8;
    sum3 = o; o += (size_t)F * 8;
    sq3 = o; o += (size_t)F * 8;
    enc64 = o; o += (size_t)B * F * 8;
    zero_end = o;
    xstat = o; o += 16 * 8;
    stats = o; o += up((size_t)kStatFloats(F) * 4, 16);
    train_par = o; o += 512 * 4;
    argmax = o; o += up((size_t)B * F * 4, 16);
    yhat3 = o; o += up((size_t)B * F * 4, 16);
    o = up(o, 1024);
    packed = o; o += 32768 + (size_t)(F / 128) * 65536;
    h2img = o;
    tiles2 = 2 * ((P + 255) / 256);
    o += (size_t)B * tiles2 * 32768;
    total = o;
  }
};

// pointnet_train_split.cu: the fp32-accurate (fp16 hi/lo split, three MMAs per product) statistics pass and fused forward
int split_pack(const float* conv2_w, const float* conv3_w, const float* bn3_w, int F, char* packed, int sms, cudaStream_t st);
int split_stats2(const float* x, int B, int P, const char* packed, const float* train_par, double* sum2, double* sq2, int sms,
                 cudaStream_t st);
int split_forward(const float* x, int B, int P, int F, const char* packed, const float* train_par, char* h2img,
                  unsigned long long* enc64, double* sum3, double* sq3, int sms, cudaStream_t st);

// pointnet_backward.cu: sum-over-ranks blocks of the backward workspace (sync points 3..5)
int backward_sync_blocks(int F, int sync_point, int* n_blocks, int* buffer, size_t* byte_offset, int64_t* count, int* is_f64);

}  // namespace pn
}  // namespace crdpn
