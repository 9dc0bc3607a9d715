// pointnet_backward.cu -- backward of the train-mode PointNet encoder (autograd of ShapeEncoderPC.forward,
// auxiliary/model.py:174-180, reached from loss.backward() at training.py:75).
//
// Structure that this backward exploits (instead of replaying the stock chain rule over the B x F x P tensor):
//   * max over points: grad of z3 = bn3(conv3(h2)) is non-zero at ONE point per (cloud, channel) -- the arg-max the
//     forward recorded.  So the F-wide layer never has to be evaluated again.
//   * train-mode BatchNorm couples all points of the batch: grad_y = (gamma/sigma) (g - mean(g) - yhat mean(g yhat)).
//     Through conv3 the two mean terms become an AFFINE function of h2:  grad_h2[n] = Sp[n] + u' + Q h2[n], with
//     Q = -(1/M) W3^T diag(a3 dgamma3 / sigma3) W3  (128 x 128), u' a 128-vector, Sp the sparse arg-max part.
//     The same happens one layer down with a 64 x 64 matrix Q1.  The weight gradients of the mean terms reduce to
//     second moments of the activations (W C, C = sum (h - m)(h - m)^T).
//   * everything downstream of grad_h2 is linear in it, so the sparse part runs through the very same kernel as a
//     second stream of B*F "virtual points" (entry (b,c) -> point argmax[b,c], gradient coef[b,c] * W3[c,:]).
// Dense work per point is therefore 128x128 + 2 * 128x64 (+ moments) instead of 2 * 128 x F: all of it on tcgen05
// tensor cores (bf16 operands, fp32 accumulation in TMEM), tile = 128 points, accumulator lanes = channels so that
// every per-channel reduction over points is a per-thread sum over TMEM columns.
//
// Launch sequence (crdpn_pointnet_backward):
//   l3_reduce (dbeta3, dgamma3) | h2_colsum (S2) | q_kernel (Q, u', W2^T image) | gs_kernel (arg-max gather for dW3)
//   | pass2 (tcgen05: streams A+B through layer 2, moments M2/M1, local part of layer 1) | reduce partials
//   | mid (dgamma2, Q1, u1') | pass3 (tcgen05: BN2's dense term through layer 1) | final (dW3, dW2, dW1, BN grads)
#include "pointnet_common.cuh"
#include "pointnet_train.cuh"
#include <stdlib.h>

namespace crdpn {
namespace pn {

constexpr int kPartFloats = 128 * 64 + 128 * 128 + 128 * 128;  // per-CTA partials: T2 | M2 | M1
constexpr int kPartT2 = 0, kPartM2 = 128 * 64, kPartM1 = 128 * 64 + 128 * 128;
constexpr int kCov2 = 0, kCov1 = 128 * 128, kCovM2 = kCov1 + 64 * 64, kCovM1 = kCovM2 + 128, kCovFloats = kCovM1 + 64;

struct BwdWs {
  size_t zero_begin, S2, acc2, acc1, zero_end;  // doubles: S2[128] | dbeta2[128], gzh2[128] | db1[64], gzh1[64], T1[64][3]
  size_t red;      // float[kPartFloats]: reduced T2 | M2 | M1
  size_t cov;      // float[kCovFloats]: Cov(h2) | Cov(h1) | mean h2 | mean h1
  size_t Gs;       // float[F*128]
  size_t vec;      // float[512]: uprime[128] | a2[128] | u1prime[64] | pad
  size_t qimg;     // bf16 [128][128] operand image of Q (32 KB)
  size_t w2timg;   // bf16 [128][128] operand image of W2^T (rows j < 64 used)
  size_t q1img;    // bf16 [128][64] operand image of Q1 (rows j < 64 used)
  size_t part;     // float[grid][kPartFloats]
  size_t total;
  __host__ __device__ BwdWs(int F, int grid) {
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t o = 0;
    zero_begin = o;
    S2 = o; o += 128 * 8;
    acc2 = o; o += 256 * 8;
    acc1 = o; o += 64 * 5 * 8;
    zero_end = o;
    red = o; o += (size_t)kPartFloats * 4;
    cov = o; o += (size_t)kCovFloats * 4;
    Gs = o; o += (size_t)F * 128 * 4;
    vec = o; o += 512 * 4;
    o = up(o, 1024);
    qimg = o; o += 32768;
    w2timg = o; o += 32768;
    q1img = o; o += 16384;
    part = o; o += (size_t)grid * kPartFloats * 4;
    total = o;
  }
};

// ---------------------------------------------------------------------------------------------------------
// dbeta3[c] = sum_b g[b,c];  dgamma3[c] = sum_b g[b,c] * yhat3[b,c]        (one warp per channel)
__global__ void __launch_bounds__(256) pn_bwd_l3_reduce_kernel(const float* __restrict__ g, const float* __restrict__ yhat3,
                                                               int B, int F, float* __restrict__ dgamma3, float* __restrict__ dbeta3) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= F) return;
  float sb = 0.f, sg = 0.f;
  for (int b = lane; b < B; b += 32) {
    const float gv = g[(size_t)b * F + c];
    sb += gv;
    sg = fmaf(gv, yhat3[(size_t)b * F + c], sg);
  }
  for (int off = 16; off >= 1; off >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sg += __shfl_xor_sync(0xffffffffu, sg, off);
  }
  if (lane == 0) { dbeta3[c] = sb; dgamma3[c] = sg; }
}

// S2[k] = sum over all real points of h2[n][k]  (bf16 tiles written by the forward)
__global__ void __launch_bounds__(256) pn_h2_colsum_kernel(const char* __restrict__ h2img, int B, int P, int tiles2, double* __restrict__ S2) {
  const int chunk = threadIdx.x & 15, rg = threadIdx.x >> 4;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int ntiles = B * tiles2;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tl = tile % tiles2;
    int valid = P - tl * 128;
    valid = valid < 0 ? 0 : (valid > 128 ? 128 : valid);
    const char* tb = h2img + (size_t)tile * kTileBytes + (chunk >> 3) * kKBlockBytes;
    for (int r = rg; r < valid; r += 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(tb + sw128_off(r, (chunk & 7) * 8));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
  __shared__ float red[16][128];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rg][chunk * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
    for (int r = 0; r < 16; ++r) s += red[r][threadIdx.x];
    atomicAdd(S2 + threadIdx.x, (double)s);
  }
}

// Q[k][j] = -(1/M) sum_c a3 dgamma3 istd3 W3[c][k] W3[c][j]  -> bf16 operand image;  u'[k] = u[k] - (Q m2)[k],
// u[k] = -(1/M) sum_c a3 dbeta3 W3[c][k].   Block 128: W2^T operand image and a2 = gamma2 * istd2.
struct QParams {
  const float *c3w, *g3, *dgamma3, *dbeta3, *stats, *c2w, *g2;
  const double* S2;
  double M;
  int F;
  char *qimg, *w2timg;
  float* vec;
};
// 512 threads: thread (part, j) sums channels [part*F/4, (part+1)*F/4) for column j with 8 loads in flight; the four
// parts fold through shared memory (the single 1024-deep dependent chain of the first version cost 58 us).
__global__ void __launch_bounds__(512) pn_bwd_q_kernel(const QParams a) {
  const int t = threadIdx.x;
  __shared__ float red[128];
  if (blockIdx.x == 128) {
    // W2T[j][k] = a2[k] W2[k][j], a2 = gamma2 * istd2 (BN2's scale on the way back); rows j >= 64 are zero
    for (int i = t; i < 128 * 128; i += 512) {
      const int j = i >> 7, k = i & 127;
      const float v = j < 64 ? a.c2w[k * 64 + j] * (a.g2[k] * a.stats[kStatIstd2 + k]) : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(a.w2timg + (k >> 6) * kKBlockBytes + sw128_off(j, k & 63)) = __float2bfloat16_rn(v);
    }
    if (t < 128) a.vec[128 + t] = a.g2[t] * a.stats[kStatIstd2 + t];
    return;
  }
  const int k = blockIdx.x, j = t & 127, part = t >> 7;
  const float* istd3 = a.stats + kStatIstd3(a.F);
  __shared__ float cqw[1024], cuw[1024];  // per channel c: coefficient * W3[c][k]
  __shared__ float qpart[4][128], upart[4][128];
  for (int c = t; c < a.F; c += 512) {
    const float a3 = a.g3[c] * istd3[c];
    const float wk = __ldg(a.c3w + c * 128 + k);
    cqw[c] = a3 * a.dgamma3[c] * istd3[c] * wk;
    cuw[c] = a3 * a.dbeta3[c] * wk;
  }
  __syncthreads();
  const int c0 = part * (a.F >> 2), c1 = c0 + (a.F >> 2);   // F is a multiple of 128
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, q4 = 0.f, q5 = 0.f, q6 = 0.f, q7 = 0.f;
  for (int c = c0; c < c1; c += 8) {
    const float w0 = __ldg(a.c3w + (c + 0) * 128 + j), w1 = __ldg(a.c3w + (c + 1) * 128 + j);
    const float w2 = __ldg(a.c3w + (c + 2) * 128 + j), w3 = __ldg(a.c3w + (c + 3) * 128 + j);
    const float w4 = __ldg(a.c3w + (c + 4) * 128 + j), w5 = __ldg(a.c3w + (c + 5) * 128 + j);
    const float w6 = __ldg(a.c3w + (c + 6) * 128 + j), w7 = __ldg(a.c3w + (c + 7) * 128 + j);
    q0 = fmaf(cqw[c + 0], w0, q0); q1 = fmaf(cqw[c + 1], w1, q1); q2 = fmaf(cqw[c + 2], w2, q2); q3 = fmaf(cqw[c + 3], w3, q3);
    q4 = fmaf(cqw[c + 4], w4, q4); q5 = fmaf(cqw[c + 5], w5, q5); q6 = fmaf(cqw[c + 6], w6, q6); q7 = fmaf(cqw[c + 7], w7, q7);
  }
  qpart[part][j] = ((q0 + q1) + (q2 + q3)) + ((q4 + q5) + (q6 + q7));
  float u = 0.f;
  for (int c = c0 + j; c < c1; c += 128) u += cuw[c];
  upart[part][j] = u;
  __syncthreads();
  if (t < 128) {
    const float q = (qpart[0][j] + qpart[1][j]) + (qpart[2][j] + qpart[3][j]);
    const float us = (upart[0][j] + upart[1][j]) + (upart[2][j] + upart[3][j]);
    const float invM = (float)(-1.0 / a.M);
    const __nv_bfloat16 qb = __float2bfloat16_rn(q * invM);
    *reinterpret_cast<__nv_bfloat16*>(a.qimg + (j >> 6) * kKBlockBytes + sw128_off(k, j & 63)) = qb;
    // u'[k] = sum_j ( u_partial[j] * invM - Q_bf16[k][j] * m2[j] )
    red[t] = us * invM - __bfloat162float(qb) * (float)(a.S2[j] / a.M);
  }
  __syncthreads();
  for (int off = 64; off >= 1; off >>= 1) {
    if (t < off) red[t] += red[t + off];
    __syncthreads();
  }
  if (t == 0) a.vec[k] = red[0];
}

// Gs[c][k] = sum_b g[b,c] * h2[argmax[b,c]][k]      (one block per channel: 8 warps split the batch, lanes over k)
__global__ void __launch_bounds__(256) pn_bwd_gs_kernel(const float* __restrict__ g, const int* __restrict__ argmax,
                                                        const char* __restrict__ h2img, int B, int F, int tiles2,
                                                        float* __restrict__ Gs) {
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = lane * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  // warp w owns clouds b = w + 8 i.  Lane i first fetches (arg-max point, gradient) of cloud i of the current group of 32,
  // then the warp walks the group four clouds at a time so that four independent row gathers are in flight
  for (int base = 0; base < B; base += 256) {
    const int bl = base + warp + 8 * lane;
    int n_l = 0;
    float g_l = 0.f;
    if (bl < B) {
      n_l = __ldg(argmax + (size_t)bl * F + c);
      g_l = __ldg(g + (size_t)bl * F + c);
    }
    const int cnt = min(32, (B - base - warp + 7) / 8);   // clouds of this warp in the group
    for (int i = 0; i < cnt; i += 4) {
      uint2 v[4];
      float gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int src = min(i + u, 31);
        const int n = __shfl_sync(0xffffffffu, n_l, src);
        gv[u] = (i + u < cnt) ? __shfl_sync(0xffffffffu, g_l, src) : 0.f;
        const int b = base + warp + 8 * min(i + u, cnt - 1);
        const char* tb = h2img + ((size_t)b * tiles2 + (n >> 7)) * kTileBytes + (k0 >> 6) * kKBlockBytes;
        v[u] = *reinterpret_cast<const uint2*>(tb + sw128_off(n & 127, k0 & 63));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0 = fmaf(gv[u], bf16_lo(v[u].x), a0); a1 = fmaf(gv[u], bf16_hi(v[u].x), a1);
        a2 = fmaf(gv[u], bf16_lo(v[u].y), a2); a3 = fmaf(gv[u], bf16_hi(v[u].y), a3);
      }
    }
  }
  __shared__ float4 red[8][32];
  red[warp][lane] = make_float4(a0, a1, a2, a3);
  __syncthreads();
  if (warp == 0) {
    float4 r = red[0][lane];
    for (int w = 1; w < 8; ++w) { const float4 v = red[w][lane]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    *reinterpret_cast<float4*>(Gs + (size_t)c * 128 + k0) = r;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pass 2 (tcgen05): one 128-point tile at a time.
//   stream A (real points):    D1[k][n] = Q[k][j] h2[n][j]^T;  grad_h2 = D1 + u'
//   stream B (virtual points): grad_h2[n] = coef[e] * W3[c_e][:],  h2 / x taken at the arg-max point of entry e
//   gz2 = grad_h2 where h2 > 0;  per-channel sums (dbeta2, sum gz2 h2);  T2[k][j] += gz2^T h1;  D3[j][n] = W2^T (a2 gz2)
//   -> gz1 = D3 where z1 > 0: db1, sum gz1 h1, T1 += gz1 x^T.   Stream A also accumulates M2 = h2^T h2, M1 = h1^T h1
//   (row 64 of the h1^T image is ones, so column 64 of M1 is S1).
struct Pass2Params {
  const float* x;
  int B, P, F;
  int tiles2, nA, nB;        // tiles per cloud, stream-A tiles, stream-B tiles
  long long E;               // B*F virtual points
  const char* h2img;
  const int* argmax;
  const float* g;            // grad_out [B,F]
  const float* c3w;          // [F][128]
  const float* g3;
  const float* stats;
  const float* train_par;    // W1p
  const char *qimg, *w2timg;
  const float* vec;          // uprime[128] | a2[128]
  double *acc2, *acc1;
  float* part;
  int debug;   // CRDPN_PN_DEBUG: block 0 prints per-phase cycle totals
};
constexpr uint32_t kP2H2 = 0, kP2Q = 32768, kP2H2T = 65536, kP2GZT = 98304, kP2H1T = 131072, kP2W2T = 163840;
constexpr uint32_t kP2X = 196608, kP2Par = kP2X + 1536, kP2Ent = kP2Par + 2048, kP2Bar = kP2Ent + 1024;
constexpr uint32_t kP2Smem = kP2Bar + 64 + 1024;

__device__ __forceinline__ uint32_t koff128(int kk) {  // descriptor offset of the kk-th K=16 step (two 64-wide K-blocks)
  return (uint32_t)(kk >> 2) * (kKBlockBytes >> 4) + (uint32_t)(kk & 3) * 2u;
}

// epilogue 1 of pass 2: thread = layer-2 channel k (TMEM lane), its 64 of the tile's 128 points (columns).
//   gz2[n][k] = grad_h2[n][k] where h2[n][k] > 0;  written three ways: GY[n][k] (bf16, in place of h2: B operand of
//   D3 = (a2 W2)^T gz2^T), GZT[k][n] (A operand of T2 += gz2^T h1), and for stream A the transposed h2 image H2T[k][n].
// A: real points (grad_h2 = Q h2 + u' from TMEM); !A: virtual points (coef[e] * W3[c_e][k]).  FULL: no padding rows.
template <bool A, bool FULL>
__device__ __forceinline__ void pass2_epilogue1(uint8_t* sm, uint32_t trow, int k, int hf, float upk, int nvalid,
                                                const int* centry, const float* coefs, const float* __restrict__ c3w) {
  uint8_t* h2row = sm + kP2H2 + (k >> 6) * kKBlockBytes;
  const int kk = k & 63;
  int w3c = -1;
  float w3v = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int nb = 64 * hf + 32 * half;
    uint32_t acc[32];
    if (A) tmem_ld32(trow + (uint32_t)nb, acc);
    // every shared-memory read of this half first: h2 is overwritten in place below
    uint32_t hb[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) hb[i] = *reinterpret_cast<const uint16_t*>(h2row + sw128_off(nb + i, kk));
    if (!FULL) {
#pragma unroll
      for (int i = 0; i < 32; ++i) hb[i] = (nb + i < nvalid) ? hb[i] : 0u;  // padding rows repeat a real point: gate them off
    }
    float v[32];
    if (A) {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) + upk;
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int ce = centry[nb + i];
        if (ce != w3c) { w3c = ce; w3v = __ldg(c3w + (size_t)ce * 128 + k); }  // warp-uniform, <= 2 per tile
        v[i] = coefs[nb + i] * w3v;
      }
    }
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const int n = nb + g8 * 8 + i;
        // h2 is a ReLU output (never negative): h2 > 0  <=>  its bf16 bits are non-zero
        const float g0 = hb[g8 * 8 + i] ? v[g8 * 8 + i] : 0.f;
        const float g1 = hb[g8 * 8 + i + 1] ? v[g8 * 8 + i + 1] : 0.f;
        const uint32_t w = pack_bf16(g0, g1);
        pk[i >> 1] = w;
        *reinterpret_cast<uint16_t*>(h2row + sw128_off(n, kk)) = (uint16_t)w;               // GY[n][k]
        *reinterpret_cast<uint16_t*>(h2row + sw128_off(n + 1, kk)) = (uint16_t)(w >> 16);   // GY[n+1][k]
      }
      const int n8 = nb + g8 * 8;
      const uint32_t off = (uint32_t)(n8 >> 6) * kKBlockBytes + sw128_off(k, n8 & 63);
      *reinterpret_cast<uint4*>(sm + kP2GZT + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (A) {
        uint4 t;
        t.x = hb[g8 * 8 + 0] | (hb[g8 * 8 + 1] << 16); t.y = hb[g8 * 8 + 2] | (hb[g8 * 8 + 3] << 16);
        t.z = hb[g8 * 8 + 4] | (hb[g8 * 8 + 5] << 16); t.w = hb[g8 * 8 + 6] | (hb[g8 * 8 + 7] << 16);
        *reinterpret_cast<uint4*>(sm + kP2H2T + off) = t;
      }
    }
  }
}

// TMEM columns of pass 2: [0,128) D1 then D3 | [128,208) T2 (64 h1 channels + the ones column = dbeta2) |
// [208,336) M2 | [336,464) M1
constexpr uint32_t kTmT2 = 128, kTmM2 = 208, kTmM1 = 336;

__global__ void __launch_bounds__(256, 1) pn_bwd_pass2_kernel(const Pass2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  // bar_mma: the accumulator the next epilogue reads (D1, then D3); bar_acc: the persistent accumulators T2 / M1 / M2,
  // whose MMAs run on under epilogue 3 and are only waited for before their operands are overwritten
  const uint32_t bar_ld = base + kP2Bar, bar_mma = base + kP2Bar + 8, bar_acc = base + kP2Bar + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kP2Bar + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;

  float* xs = reinterpret_cast<float*>(sm + kP2X);          // [3][128]
  float* par = reinterpret_cast<float*>(sm + kP2Par);        // uprime[128] | (unused)[128] | W1p[256]
  int* centry = reinterpret_cast<int*>(sm + kP2Ent);         // [128]
  float* coefs = reinterpret_cast<float*>(sm + kP2Ent + 512);  // [128]

  // one-time: operand images that do not change, parameters, zero rows of the h1^T image
  for (int i = tid; i < 2048; i += 256) {
    reinterpret_cast<uint4*>(sm + kP2Q)[i] = reinterpret_cast<const uint4*>(p.qimg)[i];
    reinterpret_cast<uint4*>(sm + kP2W2T)[i] = reinterpret_cast<const uint4*>(p.w2timg)[i];
    reinterpret_cast<uint4*>(sm + kP2H1T)[i] = make_uint4(0, 0, 0, 0);
  }
  par[tid] = p.vec[tid];
  par[256 + tid] = p.train_par[tid];
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float4* w1p = reinterpret_cast<const float4*>(par + 256);
  const float* istd3 = p.stats + kStatIstd3(p.F);
  constexpr uint32_t kI128 = make_idesc(128, 128), kI80 = make_idesc(128, 80);

  uint32_t n_ld = 0, n_mma = 0, n_acc = 0;   // completed phases of the barriers (uniform across the CTA)
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tph = clock64();
  auto mark = [&](int i) { if (p.debug) { const long long t = clock64(); ph[i] += t - tph; tph = t; } };
  int ntiles_done = 0;
  bool t2_started = false, m_started = false;
  float s_b1 = 0.f, s_t0 = 0.f, s_t1 = 0.f, s_t2 = 0.f;  // layer-1 channel j = 32q+lane (q < 2), this thread's column half
  const int k = q * 32 + lane;
  const float upk = par[k];
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
  const int ntot = p.nA + p.nB;

  auto tile_valid = [&](int tile) -> int {  // number of real rows of a tile (0: skip it)
    if (tile < p.nA) {
      const int nv = p.P - (tile % p.tiles2) * 128;
      return nv > 128 ? 128 : (nv < 0 ? 0 : nv);
    }
    const long long left = p.E - (long long)(tile - p.nA) * 128;
    return (int)(left < 128 ? left : 128);
  };
  auto next_tile = [&](int tile) -> int {
    tile += gridDim.x;
    while (tile < ntot && tile_valid(tile) == 0) tile += gridDim.x;
    return tile;
  };
  int tile = (int)blockIdx.x - (int)gridDim.x;
  tile = next_tile(tile);
  bool h2_prefetched = false, x_prefetched = false, b_prefetched = false;
  float px0 = 0.f, px1 = 0.f, px2 = 0.f;
  // stream-B gather of one virtual-point row (thread = (row r, 64-channel half)): arg-max indirection, then the h2 row,
  // the point's coordinates and the entry's coefficient, all into registers -- issued one tile ahead under epilogue 3
  uint4 brow[8];
  int bc = 0;
  float bcoef = 0.f;
  int bn = 0, bbb = 0;
  size_t bge = 0;
  auto gather_b_index = [&](int t) {   // stage 1: which point? (issued before the MMA wait so that its latency hides there)
    const int r = tid >> 1;
    if (r >= tile_valid(t)) return;
    const long long e = (long long)(t - p.nA) * 128 + r;
    bc = (int)(e / p.B);               // channel-major: a tile sees <= 2 channels
    bbb = (int)(e - (long long)bc * p.B);
    bge = (size_t)bbb * p.F + bc;
    bn = p.argmax[bge];
  };
  auto gather_b_rows = [&](int t) {    // stage 2: the h2 row, the point's coordinates, the entry's coefficient
    const int r = tid >> 1, part = tid & 1;
    if (r >= tile_valid(t)) return;
    const char* src = p.h2img + ((size_t)bbb * p.tiles2 + (bn >> 7)) * kTileBytes + part * kKBlockBytes;
    const int sr = bn & 127;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) brow[jj] = *reinterpret_cast<const uint4*>(src + sw128_off(sr, jj * 8));
    if (part == 0) {
      const float* xc = p.x + (size_t)bbb * 3 * p.P;
      px0 = __ldg(xc + bn); px1 = __ldg(xc + p.P + bn); px2 = __ldg(xc + 2 * p.P + bn);
      bcoef = p.g3[bc] * istd3[bc] * p.g[bge];
    }
  };
  auto gather_b = [&](int t) { gather_b_index(t); gather_b_rows(t); };

  while (tile < ntot) {
    const bool A = tile < p.nA;
    const int nvalid = tile_valid(tile);
    const int nxt = next_tile(tile);
    ++ntiles_done;
    // ---- load phase
    if (A) {
      const int b = tile / p.tiles2, tl = tile % p.tiles2;
      if (tid == 0 && !h2_prefetched) {
        mbar_expect_tx(bar_ld, kTileBytes);
        const char* src = p.h2img + (size_t)tile * kTileBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(base + kP2H2 + c * 8192u, src + c * 8192, 8192u, bar_ld);
      }
      if (tid < 128) {
        if (!x_prefetched) {
          const int n = tl * 128 + tid;
          const float* xc = p.x + (size_t)b * 3 * p.P;
          const bool ok = tid < nvalid;
          px0 = ok ? __ldg(xc + n) : 0.f;
          px1 = ok ? __ldg(xc + p.P + n) : 0.f;
          px2 = ok ? __ldg(xc + 2 * p.P + n) : 0.f;
        }
        xs[tid] = px0; xs[128 + tid] = px1; xs[256 + tid] = px2;
      }
    } else {
      const int r = tid >> 1, part = tid & 1;
      uint8_t* dst = sm + kP2H2 + part * kKBlockBytes;
      if (r < nvalid) {
        if (!b_prefetched) gather_b(tile);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint4*>(dst + sw128_off(r, jj * 8)) = brow[jj];
        if (part == 0) {
          xs[r] = px0; xs[128 + r] = px1; xs[256 + r] = px2;
          centry[r] = bc;
          coefs[r] = bcoef;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint4*>(dst + sw128_off(r, jj * 8)) = make_uint4(0, 0, 0, 0);
        if (part == 0) { xs[r] = 0.f; xs[128 + r] = 0.f; xs[256 + r] = 0.f; centry[r] = 0; coefs[r] = 0.f; }
      }
    }
    h2_prefetched = false;
    x_prefetched = false;
    b_prefetched = false;
    __syncthreads();
    mark(0);
    // ---- MMA phase 1 (stream A): D1 = Q h2^T is issued now and runs under the h1^T build below
    if (A && warp == 0) {
      mbar_wait(bar_ld, n_ld & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = umma_desc_sw128(base + kP2Q), dh2 = umma_desc_sw128(base + kP2H2);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_f16(tmem, dq + koff128(kk), dh2 + koff128(kk), kI128, kk > 0);
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
    // the previous tile's T2 / M1 / M2 MMAs still read H1T, GZT and H2T: they must have retired before those are rewritten
    if (ntiles_done > 1) {
      mbar_wait(bar_acc, n_acc & 1u);
      ++n_acc;
      tc_fence_after();
    }
    // ---- h1^T operand image: rows j < 64 = relu(bn1(conv1 x)) (bf16), row 64 = 1 for real points
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int item = tid + 256 * it;
      const int j = item >> 4, ch = item & 15;
      const float4 w = w1p[j];
      const float4 xa0 = *reinterpret_cast<const float4*>(xs + ch * 8), xb0 = *reinterpret_cast<const float4*>(xs + ch * 8 + 4);
      const float4 xa1 = *reinterpret_cast<const float4*>(xs + 128 + ch * 8), xb1 = *reinterpret_cast<const float4*>(xs + 128 + ch * 8 + 4);
      const float4 xa2 = *reinterpret_cast<const float4*>(xs + 256 + ch * 8), xb2 = *reinterpret_cast<const float4*>(xs + 256 + ch * 8 + 4);
      const float x0[8] = {xa0.x, xa0.y, xa0.z, xa0.w, xb0.x, xb0.y, xb0.z, xb0.w};
      const float x1[8] = {xa1.x, xa1.y, xa1.z, xa1.w, xb1.x, xb1.y, xb1.z, xb1.w};
      const float x2[8] = {xa2.x, xa2.y, xa2.z, xa2.w, xb2.x, xb2.y, xb2.z, xb2.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaf(w.x, x0[i], fmaf(w.y, x1[i], fmaf(w.z, x2[i], w.w)));
        v[i] = (ch * 8 + i < nvalid) ? z : 0.f;
      }
      uint4 o;
      o.x = pack_relu_bf16(v[0], v[1]); o.y = pack_relu_bf16(v[2], v[3]); o.z = pack_relu_bf16(v[4], v[5]); o.w = pack_relu_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(sm + kP2H1T + (ch >> 3) * kKBlockBytes + sw128_off(j, (ch & 7) * 8)) = o;
    }
    if (tid < 16) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = tid * 8 + 2 * i;
        w[i] = (n < nvalid ? 0x3f80u : 0u) | (n + 1 < nvalid ? 0x3f800000u : 0u);
      }
      *reinterpret_cast<uint4*>(sm + kP2H1T + (tid >> 3) * kKBlockBytes + sw128_off(64, (tid & 7) * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mark(1);
    if (A) {
      mbar_wait(bar_ld, n_ld & 1u);     // every thread reads h2 from shared memory in epilogue 1
      ++n_ld;
      mbar_wait(bar_mma, n_mma & 1u);   // D1
      ++n_mma;
      tc_fence_after();
    }
    mark(2);
    // ---- epilogue 1
    if (A) {
      if (nvalid == 128) pass2_epilogue1<true, true>(sm, trow, k, hf, upk, nvalid, centry, coefs, p.c3w);
      else pass2_epilogue1<true, false>(sm, trow, k, hf, upk, nvalid, centry, coefs, p.c3w);
    } else {
      pass2_epilogue1<false, true>(sm, trow, k, hf, upk, nvalid, centry, coefs, p.c3w);  // padding rows are zero rows
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mark(3);
    // ---- MMA phase 2: D3 = (a2 W2)^T gz2^T first (epilogue 3 waits for it alone), then the persistent accumulators
    //      T2 += gz2^T [h1 | 1] ; (A) M1 += h1^T h1, M2 += h2^T h2, which run on under epilogue 3
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t dgzt = umma_desc_sw128(base + kP2GZT), dh1t = umma_desc_sw128(base + kP2H1T);
        const uint64_t dh2t = umma_desc_sw128(base + kP2H2T), dw2t = umma_desc_sw128(base + kP2W2T);
        const uint64_t dgy = umma_desc_sw128(base + kP2H2);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_f16(tmem, dw2t + koff128(kk), dgy + koff128(kk), kI128, kk > 0);
        umma_commit(bar_mma);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_f16(tmem + kTmT2, dgzt + koff128(kk), dh1t + koff128(kk), kI80, (t2_started || kk > 0) ? 1u : 0u);
        if (A) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_f16(tmem + kTmM1, dh1t + koff128(kk), dh1t + koff128(kk), kI128, (m_started || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_f16(tmem + kTmM2, dh2t + koff128(kk), dh2t + koff128(kk), kI128, (m_started || kk > 0) ? 1u : 0u);
        }
        umma_commit(bar_acc);
      }
      __syncwarp();
    }
    t2_started = true;
    if (A) m_started = true;
    if (nxt >= p.nA && nxt < ntot) gather_b_index(nxt);
    mbar_wait(bar_mma, n_mma & 1u);
    ++n_mma;
    tc_fence_after();
    mark(4);
    // the h2 buffer is free again: start the next real tile's loads now, under epilogue 3
    if (nxt < p.nA) {
      if (tid == 0) {
        mbar_expect_tx(bar_ld, kTileBytes);
        const char* src = p.h2img + (size_t)nxt * kTileBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(base + kP2H2 + c * 8192u, src + c * 8192, 8192u, bar_ld);
      }
      h2_prefetched = true;
      x_prefetched = true;
      if (tid < 128) {
        const int b = nxt / p.tiles2, tl = nxt % p.tiles2;
        const int n = tl * 128 + tid;
        const float* xc = p.x + (size_t)b * 3 * p.P;
        const bool ok = tid < tile_valid(nxt);
        px0 = ok ? __ldg(xc + n) : 0.f;
        px1 = ok ? __ldg(xc + p.P + n) : 0.f;
        px2 = ok ? __ldg(xc + 2 * p.P + n) : 0.f;
      }
    } else if (nxt < ntot) {
      gather_b_rows(nxt);
      b_prefetched = true;
    }
    // ---- epilogue 3: thread = layer-1 channel j < 64, its half of the points
    if (q < 2) {
      const float4 w = w1p[k];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nb = 64 * hf + 32 * half;
        uint32_t acc[32];
        tmem_ld32(trow + (uint32_t)nb, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 32; i4 += 4) {
          const float4 a0 = *reinterpret_cast<const float4*>(xs + nb + i4);
          const float4 a1 = *reinterpret_cast<const float4*>(xs + 128 + nb + i4);
          const float4 a2 = *reinterpret_cast<const float4*>(xs + 256 + nb + i4);
          const float x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w}, x2[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float z = fmaf(w.x, x0[i], fmaf(w.y, x1[i], fmaf(w.z, x2[i], w.w)));
            // padding rows carry gz2 = 0, so D3 is already 0 there
            const float gz1 = z > 0.f ? __uint_as_float(acc[i4 + i]) : 0.f;
            s_b1 += gz1;
            s_t0 = fmaf(gz1, x0[i], s_t0); s_t1 = fmaf(gz1, x1[i], s_t1); s_t2 = fmaf(gz1, x2[i], s_t2);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mark(5);
    tile = nxt;
  }

  if (p.debug && blockIdx.x == 0 && (tid == 0 || tid == 255))
    printf("pass2 cta0 t%d tiles %d cycles: load %lld h1t %lld mma1 %lld e1 %lld mma2 %lld e3+sync %lld\n", tid, ntiles_done,
           ph[0], ph[1], ph[2], ph[3], ph[4], ph[5]);
  // ---- flush: per-CTA partials of the three persistent accumulators, per-thread channel sums
  if (ntiles_done > 0) {
    mbar_wait(bar_acc, n_acc & 1u);
    tc_fence_after();
  }
  {
    float* part = p.part + (size_t)blockIdx.x * kPartFloats;
    uint32_t r[32];
    if (t2_started) { tmem_ld32(trow + kTmT2 + 32u * hf, r); tmem_ld_wait(); }
#pragma unroll
    for (int i = 0; i < 32; ++i) part[kPartT2 + k * 64 + 32 * hf + i] = t2_started ? __uint_as_float(r[i]) : 0.f;
    if (t2_started && hf == 0) {   // column 64 of the T2 accumulator: gz2^T 1 = dbeta2
      uint32_t d;
      tmem_ld1(trow + kTmT2 + 64u, d);
      tmem_ld_wait();
      atomicAdd(p.acc2 + k, (double)__uint_as_float(d));
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int cb = 64 * hf + 32 * half;
      if (m_started) { tmem_ld32(trow + kTmM2 + (uint32_t)cb, r); tmem_ld_wait(); }
#pragma unroll
      for (int i = 0; i < 32; ++i) part[kPartM2 + k * 128 + cb + i] = m_started ? __uint_as_float(r[i]) : 0.f;
      if (m_started) { tmem_ld32(trow + kTmM1 + (uint32_t)cb, r); tmem_ld_wait(); }
#pragma unroll
      for (int i = 0; i < 32; ++i) part[kPartM1 + k * 128 + cb + i] = m_started ? __uint_as_float(r[i]) : 0.f;
    }
    if (q < 2) {
      atomicAdd(p.acc1 + k, (double)s_b1);
      atomicAdd(p.acc1 + 128 + k * 3 + 0, (double)s_t0);
      atomicAdd(p.acc1 + 128 + k * 3 + 1, (double)s_t1);
      atomicAdd(p.acc1 + 128 + k * 3 + 2, (double)s_t2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

__global__ void __launch_bounds__(256) pn_bwd_reduce_partials_kernel(const float* __restrict__ part, int nparts, float* __restrict__ red) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kPartFloats) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 3 < nparts; c += 4) {
    s0 += part[(size_t)(c + 0) * kPartFloats + i]; s1 += part[(size_t)(c + 1) * kPartFloats + i];
    s2 += part[(size_t)(c + 2) * kPartFloats + i]; s3 += part[(size_t)(c + 3) * kPartFloats + i];
  }
  for (; c < nparts; ++c) s0 += part[(size_t)c * kPartFloats + i];
  red[i] = (s0 + s1) + (s2 + s3);
}

// cov2[j][k] = M2[j][k]/M - m2[j] m2[k] (128 x 128), cov1[i][j] likewise (64 x 64, m1 = column 64 of M1), m2, m1
__global__ void __launch_bounds__(256) pn_bwd_cov_kernel(const float* __restrict__ red, const double* __restrict__ S2, double M,
                                                         float* __restrict__ cov) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float invM = (float)(1.0 / M);
  if (i < 128 * 128) {
    const int j = i >> 7, k = i & 127;
    cov[kCov2 + i] = red[kPartM2 + i] * invM - (float)(S2[j] / M) * (float)(S2[k] / M);
  } else if (i < 128 * 128 + 64 * 64) {
    const int r = i - 128 * 128;
    const int a = r >> 6, b = r & 63;
    cov[kCov1 + r] = red[kPartM1 + a * 128 + b] * invM - (red[kPartM1 + a * 128 + 64] * invM) * (red[kPartM1 + b * 128 + 64] * invM);
  } else if (i < 128 * 128 + 64 * 64 + 128) {
    const int k = i - 128 * 128 - 64 * 64;
    cov[kCovM2 + k] = (float)(S2[k] / M);
  } else if (i < 128 * 128 + 64 * 64 + 128 + 64) {
    const int j = i - 128 * 128 - 64 * 64 - 128;
    cov[kCovM1 + j] = red[kPartM1 + j * 128 + 64] * invM;
  }
}

// dgamma2 / dbeta2, then Q1[j][i] = -(1/M) sum_k a2 dgamma2 istd2 W2[k][j] W2[k][i] (operand image, rows >= 64 zero)
// and u1'[j] = -(1/M) sum_k a2 dbeta2 W2[k][j] - (Q1 m1)[j].     grid 64 (j) x 64 threads (i)
struct MidParams {
  const float *c2w, *g2, *stats;
  const double* acc2;
  const float* red;  // T2
  const float* cov;  // mean h1
  double M;
  char* q1img;
  float* vec;
  float *d_bn2_w, *d_bn2_b;
};
__global__ void __launch_bounds__(64) pn_bwd_mid_kernel(const MidParams a) {
  const int j = blockIdx.x, i = threadIdx.x;
  __shared__ float cq[128], cu[128], red[64];
  for (int k = i; k < 128; k += 64) {
    // dgamma2 = sum_n gz2 yhat2,  yhat2 = istd2 (W2[k,:] . h1[n] - mean_raw2)  =>  istd2 (W2[k,:] . T2[k,:] - mean_raw2 dbeta2)
    const float dbeta = (float)a.acc2[k];
    const float g = a.g2[k];
    const float istd = a.stats[kStatIstd2 + k];
    float wt = 0.f;
    for (int jj = 0; jj < 64; ++jj)
      wt = fmaf(__bfloat162float(__float2bfloat16_rn(a.c2w[k * 64 + jj])), a.red[kPartT2 + k * 64 + jj], wt);
    const float dgamma = istd * (wt - a.stats[kStatMean2 + k] * dbeta);
    cq[k] = g * istd * dgamma * istd;
    cu[k] = g * istd * dbeta;
    if (j == 0) { a.d_bn2_w[k] = dgamma; a.d_bn2_b[k] = dbeta; }
  }
  __syncthreads();
  float qv = 0.f, uv = 0.f;
  for (int k = 0; k < 128; ++k) {
    const float wj = a.c2w[k * 64 + j];
    qv = fmaf(cq[k] * wj, a.c2w[k * 64 + i], qv);
    if ((k & 63) == i) uv = fmaf(cu[k], wj, uv);
  }
  const float invM = (float)(-1.0 / a.M);
  const __nv_bfloat16 qb = __float2bfloat16_rn(qv * invM);
  *reinterpret_cast<__nv_bfloat16*>(a.q1img + sw128_off(j, i)) = qb;
  *reinterpret_cast<__nv_bfloat16*>(a.q1img + sw128_off(64 + j, i)) = __float2bfloat16_rn(0.f);
  const float m1 = a.cov[kCovM1 + i];
  red[i] = uv * invM - __bfloat162float(qb) * m1;
  __syncthreads();
  for (int off = 32; off >= 1; off >>= 1) {
    if (i < off) red[i] += red[i + off];
    __syncthreads();
  }
  if (i == 0) a.vec[256 + j] = red[0];
}

// Pass 3 (tcgen05, real points only): BN2's dense term through layer 1:  grad_h1[n] += u1' + Q1 h1[n]
struct Pass3Params {
  const float* x;
  int B, P, tiles2, nA;
  const float* train_par;
  const char* q1img;
  const float* vec;   // u1prime at [256, 320)
  double* acc1;
};
constexpr uint32_t kP3H1 = 0, kP3Q1 = 16384, kP3X = 32768, kP3Par = kP3X + 1536, kP3Bar = kP3Par + 2048;
constexpr uint32_t kP3Smem = kP3Bar + 64 + 1024;

__global__ void __launch_bounds__(256, 1) pn_bwd_pass3_kernel(const Pass3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar_mma = base + kP3Bar;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kP3Bar + 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  float* xs = reinterpret_cast<float*>(sm + kP3X);
  float* par = reinterpret_cast<float*>(sm + kP3Par);  // W1p[256] | u1prime[64]

  for (int i = tid; i < 1024; i += 256) reinterpret_cast<uint4*>(sm + kP3Q1)[i] = reinterpret_cast<const uint4*>(p.q1img)[i];
  par[tid] = p.train_par[tid];
  if (tid < 64) par[256 + tid] = p.vec[256 + tid];
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float4* w1p = reinterpret_cast<const float4*>(par);
  constexpr uint32_t kI128 = make_idesc(128, 128);
  uint32_t n_mma = 0;
  float s_b1 = 0.f, s_t0 = 0.f, s_t1 = 0.f, s_t2 = 0.f;
  const int j = q * 32 + lane;

  auto tile_valid = [&](int tile) -> int {
    const int nv = p.P - (tile % p.tiles2) * 128;
    return nv > 128 ? 128 : (nv < 0 ? 0 : nv);
  };
  auto next_tile = [&](int tile) -> int {
    tile += gridDim.x;
    while (tile < p.nA && tile_valid(tile) == 0) tile += gridDim.x;
    return tile;
  };
  const int row = tid & 127, cgh = tid >> 7;   // h1 image: every thread builds half of one point's 64 channels
  auto load_x = [&](int tile, float& x0, float& x1, float& x2) {
    const int b = tile / p.tiles2, tl = tile % p.tiles2;
    const int n = tl * 128 + row;
    const float* xc = p.x + (size_t)b * 3 * p.P;
    const bool ok = row < tile_valid(tile);
    x0 = ok ? __ldg(xc + n) : 0.f;
    x1 = ok ? __ldg(xc + p.P + n) : 0.f;
    x2 = ok ? __ldg(xc + 2 * p.P + n) : 0.f;
  };
  int tile = next_tile((int)blockIdx.x - (int)gridDim.x);
  float px0 = 0.f, px1 = 0.f, px2 = 0.f;
  if (tile < p.nA) load_x(tile, px0, px1, px2);

  while (tile < p.nA) {
    const int nvalid = tile_valid(tile);
    const int nxt = next_tile(tile);
    {
      const float x0 = px0, x1 = px1, x2 = px2;
      if (cgh == 0) { xs[row] = x0; xs[128 + row] = x1; xs[256 + row] = x2; }
      const bool ok = row < nvalid;
      // h1 operand image [n][i] (K-major over the 64 layer-1 channels), zero rows for padding
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int cg = cgh * 4 + c4;
        float v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 w = w1p[cg * 8 + jj];
          const float z = fmaf(w.x, x0, fmaf(w.y, x1, fmaf(w.z, x2, w.w)));
          v[jj] = ok ? z : 0.f;
        }
        uint4 o;
        o.x = pack_relu_bf16(v[0], v[1]); o.y = pack_relu_bf16(v[2], v[3]); o.z = pack_relu_bf16(v[4], v[5]); o.w = pack_relu_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(sm + kP3H1 + sw128_off(row, cg * 8)) = o;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t dq1 = umma_desc_sw128(base + kP3Q1), dh1 = umma_desc_sw128(base + kP3H1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, dq1 + 2u * kk, dh1 + 2u * kk, kI128, kk > 0);
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
    if (nxt < p.nA) load_x(nxt, px0, px1, px2);   // in flight under the MMA and the epilogue
    mbar_wait(bar_mma, n_mma & 1u);
    ++n_mma;
    tc_fence_after();
    if (q < 2) {
      const float4 w = w1p[j];
      const float u1 = par[256 + j];
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nb = 64 * hf + 32 * half;
        uint32_t acc[32];
        tmem_ld32(trow + (uint32_t)nb, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 32; i4 += 4) {
          const float4 a0 = *reinterpret_cast<const float4*>(xs + nb + i4);
          const float4 a1 = *reinterpret_cast<const float4*>(xs + 128 + nb + i4);
          const float4 a2 = *reinterpret_cast<const float4*>(xs + 256 + nb + i4);
          const float x0[4] = {a0.x, a0.y, a0.z, a0.w}, x1[4] = {a1.x, a1.y, a1.z, a1.w}, x2[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float z = fmaf(w.x, x0[i], fmaf(w.y, x1[i], fmaf(w.z, x2[i], w.w)));
            const float gz1 = (nb + i4 + i < nvalid && z > 0.f) ? __uint_as_float(acc[i4 + i]) + u1 : 0.f;
            s_b1 += gz1;
            s_t0 = fmaf(gz1, x0[i], s_t0); s_t1 = fmaf(gz1, x1[i], s_t1); s_t2 = fmaf(gz1, x2[i], s_t2);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tile = nxt;
  }
  if (q < 2) {
    atomicAdd(p.acc1 + j, (double)s_b1);
    atomicAdd(p.acc1 + 128 + j * 3 + 0, (double)s_t0);
    atomicAdd(p.acc1 + 128 + j * 3 + 1, (double)s_t1);
    atomicAdd(p.acc1 + 128 + j * 3 + 2, (double)s_t2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128u);
  }
}

// Final assembly.  blocks [0,F): dW3 row c | [F, F+128): dW2 row k | F+128: layer 1 + zero conv-bias gradients
struct FinalParams {
  const float *c1w, *c2w, *c3w, *g1, *be1, *g2, *g3;
  const float* stats;
  const double *S2, *acc2, *acc1, *xstat;
  const float *red, *cov, *Gs;
  double M;
  int F;
  const float *dgamma3, *dbeta3;   // = d_bn3_w, d_bn3_b (already written)
  const float *dgamma2;            // = d_bn2_w
  float *d_c1w, *d_c1b, *d_c2w, *d_c2b, *d_c3w, *d_c3b, *d_bn1_w, *d_bn1_b;
};
__global__ void __launch_bounds__(128) pn_bwd_final_kernel(const FinalParams a) {
  const int t = threadIdx.x;
  const double M = a.M;
  __shared__ float wrow[128];
  if ((int)blockIdx.x < a.F) {
    const int c = blockIdx.x, k = t;
    wrow[t] = a.c3w[c * 128 + t];
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
    for (int j = 0; j < 128; j += 2) {  // (W3 Cov2)[c][k]
      acc0 = fmaf(wrow[j], __ldg(a.cov + kCov2 + j * 128 + k), acc0);
      acc1 = fmaf(wrow[j + 1], __ldg(a.cov + kCov2 + (j + 1) * 128 + k), acc1);
    }
    const float istd = a.stats[kStatIstd3(a.F) + c];
    const float a3 = a.g3[c] * istd;
    a.d_c3w[c * 128 + k] = a3 * (a.Gs[(size_t)c * 128 + k] - a.dbeta3[c] * a.cov[kCovM2 + k] - a.dgamma3[c] * istd * (acc0 + acc1));
    if (t == 0) a.d_c3b[c] = 0.f;
  } else if ((int)blockIdx.x < a.F + 128) {
    const int k = blockIdx.x - a.F, j = t;
    if (t < 64) wrow[t] = a.c2w[k * 64 + t];
    __syncthreads();
    if (j < 64) {
      float acc = 0.f;
#pragma unroll 8
      for (int i = 0; i < 64; ++i) acc = fmaf(wrow[i], __ldg(a.cov + kCov1 + i * 64 + j), acc);
      const float istd = a.stats[kStatIstd2 + k];
      const float dbeta = (float)a.acc2[k];
      const float dgamma = a.dgamma2[k];
      a.d_c2w[k * 64 + j] = a.g2[k] * istd * (a.red[kPartT2 + k * 64 + j] - dbeta * a.cov[kCovM1 + j] - dgamma * istd * acc);
    }
    if (t == 0) a.d_c2b[k] = 0.f;
  } else {
    if (t < 64) {
      const int j = t;
      const float g = a.g1[j];
      const double db = a.acc1[j];
      const double istd = a.stats[kStatIstd1 + j];
      // dgamma1 = sum_n gz1 yhat1,  yhat1 = istd1 (W1[j,:] . x[n] - mean_raw1)
      double wt = 0;
      for (int e = 0; e < 3; ++e) wt += (double)a.c1w[j * 3 + e] * a.acc1[128 + j * 3 + e];
      const double dg = istd * (wt - (double)a.stats[kStatMean1 + j] * db);
      a.d_bn1_w[j] = (float)dg;
      a.d_bn1_b[j] = (float)db;
      a.d_c1b[j] = 0.f;
      const double a1 = (double)g * istd;
      for (int d = 0; d < 3; ++d) {
        double wc = 0;
        for (int e = 0; e < 3; ++e) wc += (double)a.c1w[j * 3 + e] * a.xstat[3 + e * 3 + d];
        a.d_c1w[j * 3 + d] = (float)(a1 * (a.acc1[128 + j * 3 + d] - db * a.xstat[d] - dg * istd * wc));
      }
    }
  }
}

}  // namespace pn
}  // namespace crdpn

using namespace crdpn;

static int bwd_grid(int* grid) {
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc) return rc;
  if (di.max_smem_optin < (int)pn::kP2Smem) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet_backward: not enough shared memory");
  *grid = di.sms;
  return CRDPN_OK;
}

extern "C" int crdpn_pointnet_backward_workspace_bytes(int64_t B, int64_t P, int64_t F, size_t* bytes) {
  if (!bytes || B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_backward_workspace_bytes: bad argument");
  if (!pn::pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  int grid = 0;
  int rc = bwd_grid(&grid);
  if (rc) return rc;
  *bytes = pn::BwdWs((int)F, grid).total;
  return CRDPN_OK;
}

namespace crdpn {
namespace pn {
int backward_sync_blocks(int F, int sync_point, int* n_blocks, int* buffer, size_t* byte_offset, int64_t* count, int* is_f64) {
  const BwdWs W(F, 1);  // the offsets of the reduced blocks do not depend on the grid
  auto put = [&](int i, int buf, size_t off, int64_t n, int f64) { buffer[i] = buf; byte_offset[i] = off; count[i] = n; is_f64[i] = f64; };
  switch (sync_point) {
    case 3:  // after phase 0: dgamma3 / dbeta3 partial sums, sum of h2, arg-max gather for dW3
      put(0, 1, W.S2, 128, 1); put(1, 1, W.Gs, (int64_t)F * 128, 0); put(2, 2, 0, F, 0); put(3, 3, 0, F, 0);
      *n_blocks = 4;
      break;
    case 4:  // after phase 1: layer-2 accumulators and the reduced second moments T2 | M2 | M1
      put(0, 1, W.acc2, 256, 1); put(1, 1, W.red, kPartFloats, 0);
      *n_blocks = 2;
      break;
    case 5:  // after phase 2: layer-1 accumulators
      put(0, 1, W.acc1, 64 * 5, 1);
      *n_blocks = 1;
      break;
    default: return fail(CRDPN_E_BADARG, "crdpn_pointnet_sync_blocks: sync_point must be 0..5");
  }
  return CRDPN_OK;
}
}  // namespace pn
}  // namespace crdpn

// Phases (a rank-synchronised run all-reduces the blocks of crdpn_pointnet_sync_blocks(3..5) between them):
//   0: zero, BN3 reductions, sum of h2, arg-max gather      1: Q / u', dense pass 2, partial reduction
//   2: covariances, BN2 terms, dense pass 3                  3: parameter gradients
extern "C" int crdpn_pointnet_backward_phased(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv2_w, const float* conv3_w,
    const float* bn1_w, const float* bn1_b, const float* bn2_w, const float* bn2_b,
    const float* bn3_w, const float* bn3_b,
    const float* grad_out, const void* ctx, size_t ctx_bytes,
    float* d_conv1_w, float* d_conv1_b, float* d_conv2_w, float* d_conv2_b, float* d_conv3_w, float* d_conv3_b,
    float* d_bn1_w, float* d_bn1_b, float* d_bn2_w, float* d_bn2_b, float* d_bn3_w, float* d_bn3_b,
    void* workspace, size_t workspace_bytes, int phase_begin, int phase_end, int64_t total_points, void* stream) {
  (void)bn3_b;
  if (phase_begin < 0 || phase_end > 4 || phase_begin >= phase_end || total_points < B * P)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_backward: bad phase range / total_points");
  auto on = [&](int ph) { return phase_begin <= ph && ph < phase_end; };
  if (!x || !conv1_w || !conv2_w || !conv3_w || !bn1_w || !bn1_b || !bn2_w || !bn2_b || !bn3_w || !grad_out || !ctx ||
      !d_conv1_w || !d_conv1_b || !d_conv2_w || !d_conv2_b || !d_conv3_w || !d_conv3_b || !d_bn1_w || !d_bn1_b ||
      !d_bn2_w || !d_bn2_b || !d_bn3_w || !d_bn3_b || !workspace)
    return fail(CRDPN_E_BADARG, "crdpn_pointnet_backward: null pointer");
  if (B <= 0 || P <= 0) return fail(CRDPN_E_BADARG, "crdpn_pointnet_backward: bad size");
  if (!pn::pointnet_f_ok(F)) return fail(CRDPN_E_UNSUPPORTED, "crdpn_pointnet: feature_dim must be 128, 256, 512 or 1024");
  if (((uintptr_t)ctx & 1023) || ((uintptr_t)workspace & 1023)) return fail(CRDPN_E_ALIGN, "crdpn_pointnet_backward: ctx / workspace must be 1024-byte aligned");
  const pn::TrainCtx L((int)B, (int)P, (int)F);
  if (ctx_bytes < L.total) return fail(CRDPN_E_WORKSPACE, "crdpn_pointnet_backward: ctx too small");
  int grid = 0;
  int rc = bwd_grid(&grid);
  if (rc) return rc;
  const pn::BwdWs W((int)F, grid);
  if (workspace_bytes < W.total) return fail(CRDPN_E_WORKSPACE, "crdpn_pointnet_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const char* c = (const char*)ctx;
  char* w = (char*)workspace;
  const double M = (double)total_points;  // points of ALL ranks
  const float* stats = (const float*)(c + L.stats);
  const float* train_par = (const float*)(c + L.train_par);
  const int* argmax = (const int*)(c + L.argmax);
  const float* yhat3 = (const float*)(c + L.yhat3);
  const char* h2img = c + L.h2img;
  double* S2 = (double*)(w + W.S2);
  double* acc2 = (double*)(w + W.acc2);
  double* acc1 = (double*)(w + W.acc1);
  float* red = (float*)(w + W.red);
  float* Gs = (float*)(w + W.Gs);
  float* vec = (float*)(w + W.vec);

  static bool attr_set[64] = {false};
  int device = 0;
  CRDPN_CUDA(cudaGetDevice(&device));
  if (!attr_set[device]) {
    CRDPN_CUDA(cudaFuncSetAttribute(pn::pn_bwd_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pn::kP2Smem));
    CRDPN_CUDA(cudaFuncSetAttribute(pn::pn_bwd_pass3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pn::kP3Smem));
    attr_set[device] = true;
  }

  if (on(0)) {
    CRDPN_CUDA(cudaMemsetAsync(w + W.zero_begin, 0, W.zero_end - W.zero_begin, st));
    pn::pn_bwd_l3_reduce_kernel<<<(int)((F + 7) / 8), 256, 0, st>>>(grad_out, yhat3, (int)B, (int)F, d_bn3_w, d_bn3_b);
    CRDPN_LAUNCH_CHECK("pn_bwd_l3_reduce_kernel");
    pn::pn_h2_colsum_kernel<<<grid * 4, 256, 0, st>>>(h2img, (int)B, (int)P, L.tiles2, S2);
    CRDPN_LAUNCH_CHECK("pn_h2_colsum_kernel");
    pn::pn_bwd_gs_kernel<<<(int)F, 256, 0, st>>>(grad_out, argmax, h2img, (int)B, (int)F, L.tiles2, Gs);
    CRDPN_LAUNCH_CHECK("pn_bwd_gs_kernel");
  }
  if (on(1)) {
    pn::QParams qp{conv3_w, bn3_w, d_bn3_w, d_bn3_b, stats, conv2_w, bn2_w, S2, M, (int)F, w + W.qimg, w + W.w2timg, vec};
    pn::pn_bwd_q_kernel<<<129, 512, 0, st>>>(qp);
    CRDPN_LAUNCH_CHECK("pn_bwd_q_kernel");
  }

  pn::Pass2Params p2;
  p2.x = x; p2.B = (int)B; p2.P = (int)P; p2.F = (int)F;
  p2.tiles2 = L.tiles2; p2.nA = (int)B * L.tiles2;
  p2.E = (long long)B * F;
  p2.nB = (int)((p2.E + 127) / 128);
  p2.h2img = h2img; p2.argmax = argmax; p2.g = grad_out; p2.c3w = conv3_w; p2.g3 = bn3_w; p2.stats = stats;
  p2.train_par = train_par; p2.qimg = w + W.qimg; p2.w2timg = w + W.w2timg; p2.vec = vec;
  p2.acc2 = acc2; p2.acc1 = acc1; p2.part = (float*)(w + W.part);
  p2.debug = getenv("CRDPN_PN_DEBUG") ? 1 : 0;
  if (on(1)) {
    pn::pn_bwd_pass2_kernel<<<grid, 256, pn::kP2Smem, st>>>(p2);
    CRDPN_LAUNCH_CHECK("pn_bwd_pass2_kernel");
    pn::pn_bwd_reduce_partials_kernel<<<(pn::kPartFloats + 255) / 256, 256, 0, st>>>((const float*)(w + W.part), grid, red);
    CRDPN_LAUNCH_CHECK("pn_bwd_reduce_partials_kernel");
  }
  if (on(2)) {
    pn::pn_bwd_cov_kernel<<<(128 * 128 + 64 * 64 + 192 + 255) / 256, 256, 0, st>>>(red, S2, M, (float*)(w + W.cov));
    CRDPN_LAUNCH_CHECK("pn_bwd_cov_kernel");
    pn::MidParams mp{conv2_w, bn2_w, stats, acc2, red, (const float*)(w + W.cov), M, w + W.q1img, vec, d_bn2_w, d_bn2_b};
    pn::pn_bwd_mid_kernel<<<64, 64, 0, st>>>(mp);
    CRDPN_LAUNCH_CHECK("pn_bwd_mid_kernel");
    pn::Pass3Params p3{x, (int)B, (int)P, L.tiles2, p2.nA, train_par, w + W.q1img, vec, acc1};
    pn::pn_bwd_pass3_kernel<<<grid, 256, pn::kP3Smem, st>>>(p3);
    CRDPN_LAUNCH_CHECK("pn_bwd_pass3_kernel");
  }
  if (!on(3)) return CRDPN_OK;
  pn::FinalParams fp{conv1_w, conv2_w, conv3_w, bn1_w, bn1_b, bn2_w, bn3_w, stats, S2, acc2, acc1, (const double*)(c + L.xstat),
                     red, (const float*)(w + W.cov), Gs, M, (int)F, d_bn3_w, d_bn3_b, d_bn2_w,
                     d_conv1_w, d_conv1_b, d_conv2_w, d_conv2_b, d_conv3_w, d_conv3_b, d_bn1_w, d_bn1_b};
  pn::pn_bwd_final_kernel<<<(int)F + 129, 128, 0, st>>>(fp);
  CRDPN_LAUNCH_CHECK("pn_bwd_final_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_pointnet_backward(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv2_w, const float* conv3_w,
    const float* bn1_w, const float* bn1_b, const float* bn2_w, const float* bn2_b,
    const float* bn3_w, const float* bn3_b,
    const float* grad_out, const void* ctx, size_t ctx_bytes,
    float* d_conv1_w, float* d_conv1_b, float* d_conv2_w, float* d_conv2_b, float* d_conv3_w, float* d_conv3_b,
    float* d_bn1_w, float* d_bn1_b, float* d_bn2_w, float* d_bn2_b, float* d_bn3_w, float* d_bn3_b,
    void* workspace, size_t workspace_bytes, void* stream) {
  return crdpn_pointnet_backward_phased(x, B, P, F, conv1_w, conv2_w, conv3_w, bn1_w, bn1_b, bn2_w, bn2_b, bn3_w, bn3_b, grad_out,
                                        ctx, ctx_bytes, d_conv1_w, d_conv1_b, d_conv2_w, d_conv2_b, d_conv3_w, d_conv3_b,
                                        d_bn1_w, d_bn1_b, d_bn2_w, d_bn2_b, d_bn3_w, d_bn3_b, workspace, workspace_bytes, 0, 4,
                                        B * P, stream);
}
