// kd_losses.cu -- the reference's in-batch contrastive KD losses and its KD loss mixer, each as one or two launches
// with closed-form gradients (the eager formulation costs ~50 forward + ~80 backward launches on [138, 24] tensors).
//
//   nce_kd   : infoNCE_KD / poseNCE_KD, auxiliary/model_utils.py:225-285 (+ rotation_err, auxiliary/utils.py:156-202), and the
//              other in-batch variants of the same file: infoNCE / poseNCE (:169-223, negatives = the anchors' own rows),
//              singleinfoNCE_KD (:288-304, positive term only), multiposeNCE_KD (:307-351, every sample within 30 degrees
//              of the anchor's pose counts as a positive) -- mode bits on top of the weighting code, same three kernels
//              dropout(p) on the teacher side -> L2 normalise both -> B x B logits / tau -> pose weights -> -log(pos / sum)
//   kd_mix   : CELoss x3 + DeltaLoss (auxiliary/loss.py:7-34), TemperatureScaledKLDivLoss x7 and the weighted sum of
//              calculate_kd_loss_new (KD/vision/vanilla/vanilla_kd.py:8-32, 143-164); call site
//              KD/common/base_class.py:365-387.
//
// All of it is latency-bound ([B, 200] features, [B, 24] logits, B = 46..138): the design goal is launch count and
// determinism (fixed-order reductions, no float atomics), not bandwidth.
#include "common.cuh"

namespace crdpn {
namespace kdl {

constexpr int kNceThreads = 256;
constexpr int kMixThreads = 128;
constexpr float kPi = 3.14159265358979323846f;
// mode bits carried in the `weighting` argument above the weighting code (bits 0-2)
constexpr int kNceSelf = 0x100;     // negatives are the anchors' own normalised rows, k = n excluded (infoNCE, poseNCE)
constexpr int kNceSingle = 0x200;   // loss_n = -s_nn: the positive logit alone (singleinfoNCE_KD)
constexpr int kNceMulti = 0x400;    // positives = {k : k = n or rotation_err(n, k) <= 30 degrees} (multiposeNCE_KD)
constexpr float kMultiThresholdDeg = 30.f;

// fixed-order block sum / max (every thread returns the result)
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
__device__ __forceinline__ float block_max(float v, float* red) {
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = red[0];
  for (int w = 1; w < nw; ++w) s = fmaxf(s, red[w]);
  return s;
}

// keep / drop of element e in the dropout stream (seed, offset): word e%4 of block offset + e/4; keep iff u >= p
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long offset, long long e, float p) {
  unsigned r[4];
  philox4x32_10(seed, offset + (unsigned long long)(e >> 2), r);
  const unsigned w = (e & 3) == 0 ? r[0] : (e & 3) == 1 ? r[1] : (e & 3) == 2 ? r[2] : r[3];
  return (float)(w >> 8) * 5.9604644775390625e-08f >= p;
}

struct NceWs {  // views into the caller's workspace
  float* a_hat;   // [B, C]
  float* p_hat;   // [B, C]
  float* inv_na;  // [B]
  float* inv_np;  // [B]
  float* R;       // [B, 9]
  float* w;       // [B, B]   E_nk / S_n
  float* pfrac;   // [B]      l_pos / S_n
  float* loss_n;  // [B]
  unsigned* ticket;
};

__host__ __device__ inline size_t nce_ws_floats(long long B, long long C) { return (size_t)(2 * B * C + 2 * B + 9 * B + B * B + 2 * B + 4); }

__host__ inline NceWs nce_ws_views(void* ws, long long B, long long C) {
  NceWs v;
  float* f = (float*)ws;
  v.ticket = (unsigned*)f; f += 4;
  v.a_hat = f; f += B * C;
  v.p_hat = f; f += B * C;
  v.inv_na = f; f += B;
  v.inv_np = f; f += B;
  v.R = f; f += 9 * B;
  v.w = f; f += B * B;
  v.pfrac = f; f += B;
  v.loss_n = f;
  return v;
}

// grid B: dropout + L2 normalise row n of both sides; rotation matrix of label n (utils.py:156-178, fp32 as there)
__global__ void __launch_bounds__(kNceThreads) nce_prep_kernel(const float* __restrict__ ori, const float* __restrict__ pos,
                                                              const float* __restrict__ label, int C, float p_drop,
                                                              unsigned long long seed, unsigned long long offset, NceWs ws, int mode) {
  __shared__ float red[kNceThreads / 32];
  const int n = blockIdx.x;
  const size_t o = (size_t)n * C;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  float sa = 0.f, sp = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = ori[o + c];
    float p = pos[o + c];
    if (p_drop > 0.f) p = dropout_keep(seed, offset, (long long)o + c, p_drop) ? p * keep_scale : 0.f;
    ws.p_hat[o + c] = p;
    sa = fmaf(a, a, sa);
    sp = fmaf(p, p, sp);
  }
  sa = block_sum(sa, red);
  sp = block_sum(sp, red);
  const float ia = 1.0f / fmaxf(sqrtf(sa), 1e-12f), ip = 1.0f / fmaxf(sqrtf(sp), 1e-12f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    ws.a_hat[o + c] = ori[o + c] * ia;
    ws.p_hat[o + c] *= ip;
  }
  if (threadIdx.x == 0) {
    ws.inv_na[n] = ia;
    ws.inv_np[n] = ip;
    if (n == 0) { ws.ticket[0] = 0u; ws.ticket[1] = (unsigned)mode; }   // the backward reads the mode from here
    if (label != nullptr) {
      const float azi = label[3 * n] * kPi / 180.f, ele = (label[3 * n + 1] - 180.f) * kPi / 180.f,
                  rol = (label[3 * n + 2] - 180.f) * kPi / 180.f;
      const float ca = cosf(azi), sa_ = sinf(azi), ce = cosf(ele), se = sinf(ele), cr = cosf(rol), sr = sinf(rol);
      float* R = ws.R + 9 * n;
      R[0] = cr * ca - sr * ce * sa_;  R[1] = sr * ca + cr * ce * sa_;   R[2] = se * sa_;
      R[3] = -cr * sa_ - sr * ce * ca; R[4] = -sr * sa_ + cr * ce * ca;  R[5] = se * ca;
      R[6] = sr * se;                  R[7] = -cr * se;                  R[8] = ce;
    }
  }
}

__device__ __forceinline__ float pose_weight(const float* Rn, const float* Rk, int weighting) {
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) t = fmaf(Rn[i], Rk[i], t);
  t = fminf(fmaxf(t, -1.f), 3.f);
  const float d = acosf((t - 1.f) * 0.5f) * (180.f / kPi) / 180.f;  // [0, 1]
  switch (weighting) {
    case 1: return d;
    case 2: return d * d;
    case 3: return sqrtf(d);
    case 4: return fabsf(sinf(d * kPi));
    default: { const float s = sinf(d * kPi); return s * s; }
  }
}

__device__ __forceinline__ float pose_dist_deg(const float* Rn, const float* Rk) {
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) t = fmaf(Rn[i], Rk[i], t);
  t = fminf(fmaxf(t, -1.f), 3.f);
  return acosf((t - 1.f) * 0.5f) * (180.f / kPi);
}

// grid B, dynamic smem (C + B) floats: row n of the logits, its soft weights (= d loss_n / d s_nk), its loss term; the
// last block reduces the loss.  `weighting` = weighting code | mode bits (kNceSelf / kNceSingle / kNceMulti).
__global__ void __launch_bounds__(kNceThreads) nce_rows_kernel(int B, int C, float inv_tau, int weighting, NceWs ws,
                                                              float* __restrict__ loss_out) {
  extern __shared__ float smem[];
  __shared__ float red[kNceThreads / 32];
  __shared__ float s_pos_sh;
  __shared__ bool last;
  float* s_a = smem;       // [C]
  float* s_e = smem + C;   // [B]
  const int mode = weighting & ~0xff, wcode = weighting & 0xff;
  const bool self = (mode & kNceSelf) != 0, single = (mode & kNceSingle) != 0, multi = (mode & kNceMulti) != 0;
  const float* neg = self ? ws.a_hat : ws.p_hat;
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_a[c] = ws.a_hat[(size_t)n * C + c];
  __syncthreads();
  for (int k = warp; k < B + 1; k += nw) {   // k = B: the positive logit a_n . p_n (its own slot: the self modes need it)
    const float* pk = k < B ? neg + (size_t)k * C : ws.p_hat + (size_t)n * C;
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(s_a[c], pk[c], d);
    for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (lane == 0) {
      if (k < B) s_e[k] = d * inv_tau; else s_pos_sh = d * inv_tau;
    }
  }
  __syncthreads();
  const float s_nn = s_pos_sh;
  float part = 0.f, lpos, S, loss_n;
  if (single) {
    for (int k = threadIdx.x; k < B; k += blockDim.x) ws.w[(size_t)n * B + k] = 0.f;
    lpos = 0.f; S = 1.f; loss_n = -s_nn;          // d loss / d s_nn = pfrac - 1 = -1
  } else {
    float m = s_nn;
    for (int k = threadIdx.x; k < B; k += blockDim.x) m = fmaxf(m, s_e[k]);
    m = block_max(m, red);
    __syncthreads();
    if (multi) {
      float partp = 0.f;
      for (int k = threadIdx.x; k < B; k += blockDim.x) {
        const bool mark = (k == n) || pose_dist_deg(ws.R + 9 * n, ws.R + 9 * k) <= kMultiThresholdDeg;
        const float e = expf(s_e[k] - m);
        s_e[k] = mark ? -e : e;     // sign bit carries the mark to the weight pass
        part += mark ? 2.f * e : e;
        partp += mark ? e : 0.f;
      }
      S = block_sum(part, red);
      const float Lp = block_sum(partp, red);
      const float invS = 1.0f / S, invLp = 1.0f / Lp;
      for (int k = threadIdx.x; k < B; k += blockDim.x) {
        const float e = fabsf(s_e[k]);
        ws.w[(size_t)n * B + k] = s_e[k] < 0.f ? e * (2.f * invS - invLp) : e * invS;
      }
      lpos = S;                    // pfrac = 1: no separate positive slot
      loss_n = logf(S) - logf(Lp);
    } else {
      for (int k = threadIdx.x; k < B; k += blockDim.x) {
        float w = 1.0f;
        if (wcode != 0) w = (k == n) ? 0.f : pose_weight(ws.R + 9 * n, ws.R + 9 * k, wcode);
        else if (self) w = (k == n) ? 0.f : 1.f;
        const float e = expf(s_e[k] - m) * w;
        s_e[k] = e;
        part += e;
      }
      lpos = expf(s_nn - m);
      S = block_sum(part, red) + lpos;
      const float invS = 1.0f / S;
      for (int k = threadIdx.x; k < B; k += blockDim.x) ws.w[(size_t)n * B + k] = s_e[k] * invS;
      loss_n = logf(S) - (s_nn - m);
    }
  }
  if (threadIdx.x == 0) {
    ws.pfrac[n] = lpos / S;
    ws.loss_n[n] = loss_n;
    __threadfence();
    last = atomicAdd(ws.ticket, 1u) == (unsigned)(B - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float acc = 0.f;
    for (int k = threadIdx.x; k < B; k += blockDim.x) acc += __ldcg(ws.loss_n + k);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      *loss_out = acc / (float)B;
      *ws.ticket = 0u;
    }
  }
}

// grid B, dynamic smem (2B + 2C) floats: gradients of row n of BOTH inputs
__global__ void __launch_bounds__(kNceThreads) nce_grads_kernel(int B, int C, float inv_tau, float p_drop, unsigned long long seed,
                                                               unsigned long long offset, NceWs ws,
                                                               const float* __restrict__ grad_loss, float* __restrict__ d_ori,
                                                               float* __restrict__ d_pos) {
  extern __shared__ float smem[];
  __shared__ float red[kNceThreads / 32];
  float* w_row = smem;          // [B]  w[n, :]
  float* w_col = smem + B;      // [B]  w[:, n]
  float* g_a = smem + 2 * B;    // [C]
  float* g_p = g_a + C;         // [C]
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    w_row[k] = ws.w[(size_t)n * B + k];
    w_col[k] = ws.w[(size_t)k * B + n];
  }
  __syncthreads();
  const float pf = ws.pfrac[n] - 1.0f;
  const size_t o = (size_t)n * C;
  const bool self = (ws.ticket[1] & (unsigned)kNceSelf) != 0u;   // negatives were the anchors' own rows
  float da = 0.f, dp = 0.f;
  if (self) {
    // s_nk = a_n . a_k: row n of the anchors collects both roles, (w[n,k] + w[k,n]) a_k; the positive pairs a_n with p_n
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float ga0 = 0.f, ga1 = 0.f;
      int k = 0;
      for (; k + 1 < B; k += 2) {
        ga0 = fmaf(w_row[k] + w_col[k], ws.a_hat[(size_t)k * C + c], ga0);
        ga1 = fmaf(w_row[k + 1] + w_col[k + 1], ws.a_hat[(size_t)(k + 1) * C + c], ga1);
      }
      if (k < B) ga0 = fmaf(w_row[k] + w_col[k], ws.a_hat[(size_t)k * C + c], ga0);
      const float an = ws.a_hat[o + c], pn = ws.p_hat[o + c];
      const float ga = (ga0 + ga1) + pf * pn, gp = pf * an;
      g_a[c] = ga;
      g_p[c] = gp;
      da = fmaf(ga, an, da);
      dp = fmaf(gp, pn, dp);
    }
  }
  for (int c = threadIdx.x; c < C && !self; c += blockDim.x) {
    float ga0 = 0.f, ga1 = 0.f, gp0 = 0.f, gp1 = 0.f;
    int k = 0;
    for (; k + 1 < B; k += 2) {
      ga0 = fmaf(w_row[k], ws.p_hat[(size_t)k * C + c], ga0);
      ga1 = fmaf(w_row[k + 1], ws.p_hat[(size_t)(k + 1) * C + c], ga1);
      gp0 = fmaf(w_col[k], ws.a_hat[(size_t)k * C + c], gp0);
      gp1 = fmaf(w_col[k + 1], ws.a_hat[(size_t)(k + 1) * C + c], gp1);
    }
    if (k < B) {
      ga0 = fmaf(w_row[k], ws.p_hat[(size_t)k * C + c], ga0);
      gp0 = fmaf(w_col[k], ws.a_hat[(size_t)k * C + c], gp0);
    }
    const float an = ws.a_hat[o + c], pn = ws.p_hat[o + c];
    const float ga = (ga0 + ga1) + pf * pn, gp = (gp0 + gp1) + pf * an;
    g_a[c] = ga;
    g_p[c] = gp;
    da = fmaf(ga, an, da);
    dp = fmaf(gp, pn, dp);
  }
  da = block_sum(da, red);
  dp = block_sum(dp, red);
  const float scale = (grad_loss ? *grad_loss : 1.0f) * inv_tau / (float)B;
  const float sa = scale * ws.inv_na[n], sp = scale * ws.inv_np[n] * (p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    d_ori[o + c] = (g_a[c] - ws.a_hat[o + c] * da) * sa;
    if (d_pos != nullptr) {
      const bool keep = p_drop > 0.f ? dropout_keep(seed, offset, (long long)o + c, p_drop) : true;
      d_pos[o + c] = keep ? (g_p[c] - ws.p_hat[o + c] * dp) * sp : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// KD loss mixer: one block per sample row, one warp per term group
//   groups 0..2 : KL(student head i, teacher head i) + CE(head i, label i // ce_bin i)
//   groups 3..5 : KL(student head i, teacher head i) + SmoothL1 delta term of angle i-3
//   group  6    : KL(student features, teacher features)
struct MixParams {
  const float* s_out[6];
  const float* t_out[6];
  int width[6];
  const float* s_feat;
  const float* t_feat;
  int feat_dim;
  const float* label;  // [n, label_stride] float32 degrees
  int label_stride;
  int n;
  int ce_bin[3];
  int delta_bin;
  unsigned terms;      // bit i (0..5): KL head i; bit 6: KL features; bit 7+i (i<3): CE head i; bit 10: delta
  float inv_T, T, w_kl, w_rep, w_gt;
  const float* grad_loss;
  float* d_s_out[6];
  float* d_t_out[6];
  float* d_s_feat;
  float* d_t_feat;
  float* row_loss;     // [n]
  unsigned* ticket;
  float* loss_out;
};

__device__ __forceinline__ float warp_sum(float v) {
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// log-sum-exp of x[0..W) * scale over a warp
__device__ __forceinline__ float warp_lse(const float* x, int W, float scale, int lane) {
  float m = -INFINITY;
  for (int j = lane; j < W; j += 32) m = fmaxf(m, x[j] * scale);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < W; j += 32) s += expf(x[j] * scale - m);
  return m + logf(warp_sum(s));
}

template <bool BWD>
__global__ void __launch_bounds__(kMixThreads) kd_mix_kernel(const MixParams p) {
  __shared__ float term[8];
  __shared__ float red[kMixThreads / 32];
  __shared__ bool last;
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float inv_n = 1.0f / (float)p.n;
  const float up = BWD ? (p.grad_loss ? *p.grad_loss : 1.0f) : 0.f;
  if (threadIdx.x < 8) term[threadIdx.x] = 0.f;
  __syncthreads();
  for (int g = warp; g < 7; g += nw) {
    const bool is_feat = g == 6;
    const bool kl_on = (p.terms >> g) & 1u;
    const bool ce_on = g < 3 && ((p.terms >> (7 + g)) & 1u);
    const bool dl_on = g >= 3 && g < 6 && ((p.terms >> 10) & 1u);
    if (!kl_on && !ce_on && !dl_on) continue;
    const int W = is_feat ? p.feat_dim : p.width[g];
    const float* s = (is_feat ? p.s_feat : p.s_out[g]) + (size_t)row * W;
    const float* t = kl_on ? (is_feat ? p.t_feat : p.t_out[g]) + (size_t)row * W : nullptr;
    float* ds = BWD ? (is_feat ? p.d_s_feat : p.d_s_out[g]) : nullptr;
    float* dt = BWD ? (is_feat ? p.d_t_feat : p.d_t_out[g]) : nullptr;
    if (ds) ds += (size_t)row * W;
    if (dt) dt += (size_t)row * W;
    const float wk = is_feat ? p.w_rep : p.w_kl;
    float val = 0.f;
    float lse_s = 0.f, lse_t = 0.f, kl = 0.f;
    if (kl_on) {
      lse_s = warp_lse(s, W, p.inv_T, lane);
      lse_t = warp_lse(t, W, p.inv_T, lane);
      for (int j = lane; j < W; j += 32) {
        const float lq = t[j] * p.inv_T - lse_t, lp = s[j] * p.inv_T - lse_s;
        kl += expf(lq) * (lq - lp);
      }
      kl = warp_sum(kl);
      val += wk * p.T * p.T * kl * inv_n;
    }
    float lse1 = 0.f;
    int bin = -1;
    float dl_grad = 0.f;  // d loss / d s[bin] of the delta term
    if (ce_on) {
      lse1 = warp_lse(s, W, 1.0f, lane);
      bin = min(max((int)floorf(p.label[(size_t)row * p.label_stride + g] / (float)p.ce_bin[g]), 0), W - 1);
      val += p.w_gt * (lse1 - s[bin]) * inv_n;
    }
    if (dl_on) {
      const float lab = p.label[(size_t)row * p.label_stride + (g - 3)];
      const float fb = (float)p.delta_bin;
      bin = min(max((int)floorf(lab / fb), 0), W - 1);
      const float td = fmodf(lab, fb) / fb - 0.5f;
      const float th = tanhf(s[bin]);
      const float diff = 5.0f * (th * 0.5f) - 5.0f * td;
      const float ad = fabsf(diff);
      val += p.w_gt * (ad < 1.0f ? 0.5f * diff * diff : ad - 0.5f) * inv_n * (1.0f / 3.0f);
      dl_grad = p.w_gt * (ad < 1.0f ? diff : (diff > 0.f ? 1.0f : -1.0f)) * 2.5f * (1.0f - th * th) * inv_n * (1.0f / 3.0f);
    }
    if (lane == 0) term[g] = val;
    if (BWD) {
      for (int j = lane; j < W; j += 32) {
        float gs = 0.f;
        if (kl_on) {
          const float lq = t[j] * p.inv_T - lse_t, lp = s[j] * p.inv_T - lse_s;
          const float q = expf(lq);
          gs += wk * p.T * (expf(lp) - q) * inv_n;
          if (dt) dt[j] = up * wk * p.T * q * ((lq - lp) - kl) * inv_n;
        }
        if (ce_on) gs += p.w_gt * (expf(s[j] - lse1) - (j == bin ? 1.0f : 0.f)) * inv_n;
        if (dl_on && j == bin) gs += dl_grad;
        if (ds) ds[j] = up * gs;
      }
    }
  }
  if (BWD) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int g = 0; g < 7; ++g) acc += term[g];
    p.row_loss[row] = acc;
    __threadfence();
    last = atomicAdd(p.ticket, 1u) == (unsigned)(p.n - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float acc = 0.f;
    for (int k = threadIdx.x; k < p.n; k += blockDim.x) acc += __ldcg(p.row_loss + k);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      *p.loss_out = acc;
      *p.ticket = 0u;
    }
  }
}

}  // namespace kdl
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_nce_kd_workspace_bytes(int64_t B, int64_t C, size_t* bytes) {
  if (!bytes || B <= 0 || C <= 0) return fail(CRDPN_E_BADARG, "crdpn_nce_kd_workspace_bytes: bad argument");
  *bytes = kdl::nce_ws_floats(B, C) * sizeof(float);
  return CRDPN_OK;
}

static int nce_check(int64_t B, int64_t C, float tau, int weighting, float p, const void* ws, size_t ws_bytes) {
  if (B <= 0 || C <= 0 || !(tau > 0.f) || weighting < 0 || weighting > 5 || !(p >= 0.f) || !(p < 1.f))
    return fail(CRDPN_E_BADARG, "crdpn_nce_kd: bad argument");
  if (B > 8192 || C > 8192 || (size_t)(2 * B + 2 * C) * sizeof(float) > 200 * 1024)
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_nce_kd: batch / feature size beyond the shared-memory row buffers");
  if (!ws || ws_bytes < kdl::nce_ws_floats(B, C) * sizeof(float)) return fail(CRDPN_E_WORKSPACE, "crdpn_nce_kd: workspace too small");
  if ((uintptr_t)ws & 15) return fail(CRDPN_E_ALIGN, "crdpn_nce_kd: workspace must be 16-byte aligned");
  return CRDPN_OK;
}

extern "C" int crdpn_nce_kd_forward(const float* feat_ori, const float* feat_pos, const float* label, int64_t B, int64_t C,
                                    float tau, int weighting, float dropout_p, uint64_t seed, uint64_t offset, float* loss,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!feat_ori || !feat_pos || !loss) return fail(CRDPN_E_BADARG, "crdpn_nce_kd_forward: null pointer");
  if (((weighting & 0xff) != 0 || (weighting & kdl::kNceMulti)) && !label)
    return fail(CRDPN_E_BADARG, "crdpn_nce_kd_forward: pose weighting / multi-positive mode needs labels");
  const int mode = weighting & ~0xff;
  if (mode & ~(kdl::kNceSelf | kdl::kNceSingle | kdl::kNceMulti)) return fail(CRDPN_E_BADARG, "crdpn_nce_kd_forward: unknown mode bits");
  int rc = nce_check(B, C, tau, weighting & 0xff, dropout_p, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  kdl::NceWs ws = kdl::nce_ws_views(workspace, B, C);
  kdl::nce_prep_kernel<<<(unsigned)B, kdl::kNceThreads, 0, st>>>(feat_ori, feat_pos, label, (int)C, dropout_p, seed, offset, ws, mode);
  CRDPN_LAUNCH_CHECK("nce_prep_kernel");
  const size_t smem = (size_t)(B + C) * sizeof(float);
  if (smem > 48 * 1024) CRDPN_CUDA(cudaFuncSetAttribute(kdl::nce_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kdl::nce_rows_kernel<<<(unsigned)B, kdl::kNceThreads, smem, st>>>((int)B, (int)C, 1.0f / tau, weighting, ws, loss);
  CRDPN_LAUNCH_CHECK("nce_rows_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_nce_kd_backward(const float* grad_loss, int64_t B, int64_t C, float tau, float dropout_p, uint64_t seed,
                                     uint64_t offset, const void* workspace, size_t workspace_bytes, float* d_ori, float* d_pos,
                                     void* stream) {
  if (!d_ori) return fail(CRDPN_E_BADARG, "crdpn_nce_kd_backward: null pointer");
  int rc = nce_check(B, C, tau, 0, dropout_p, workspace, workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  kdl::NceWs ws = kdl::nce_ws_views(const_cast<void*>(workspace), B, C);
  const size_t smem = (size_t)(2 * B + 2 * C) * sizeof(float);
  if (smem > 48 * 1024) CRDPN_CUDA(cudaFuncSetAttribute(kdl::nce_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kdl::nce_grads_kernel<<<(unsigned)B, kdl::kNceThreads, smem, st>>>((int)B, (int)C, 1.0f / tau, dropout_p, seed, offset, ws,
                                                                   grad_loss, d_ori, d_pos);
  CRDPN_LAUNCH_CHECK("nce_grads_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_kd_mix_workspace_bytes(int64_t n, size_t* bytes) {
  if (!bytes || n <= 0) return fail(CRDPN_E_BADARG, "crdpn_kd_mix_workspace_bytes: bad argument");
  *bytes = (size_t)(n + 4) * sizeof(float);
  return CRDPN_OK;
}

static int mix_fill(kdl::MixParams* p, const float* const* s_out, const float* const* t_out, const int32_t* widths,
                    const float* s_feat, const float* t_feat, int64_t feat_dim, const float* label, int64_t label_stride,
                    int64_t n, const int32_t* ce_bin, int32_t delta_bin, uint32_t terms, float T, float w_kl, float w_rep,
                    float w_gt, void* ws, size_t ws_bytes) {
  if (n <= 0 || n >= (1ll << 31) || !(T > 0.f) || !ws) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: bad argument");
  if (ws_bytes < (size_t)(n + 4) * sizeof(float)) return fail(CRDPN_E_WORKSPACE, "crdpn_kd_mix: workspace too small");
  if (terms == 0 || terms >= (1u << 11)) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: no / unknown terms selected");
  const bool need_label = (terms >> 7) != 0;
  if (need_label && (!label || label_stride <= 0)) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: CE / delta terms need labels");
  for (int i = 0; i < 6; ++i) {
    const bool kl = (terms >> i) & 1u, ce = i < 3 && ((terms >> (7 + i)) & 1u), dl = i >= 3 && ((terms >> 10) & 1u);
    p->s_out[i] = nullptr; p->t_out[i] = nullptr; p->width[i] = 0;
    if (!(kl || ce || dl)) continue;
    if (!s_out || !widths || !s_out[i] || widths[i] <= 0) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: missing student head output");
    if (kl && (!t_out || !t_out[i])) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: missing teacher head output");
    p->s_out[i] = s_out[i];
    p->t_out[i] = kl ? t_out[i] : nullptr;
    p->width[i] = widths[i];
    if (ce && (!ce_bin || ce_bin[i] <= 0)) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: bad CE bin size");
    if (dl && delta_bin <= 0) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: bad delta bin size");
  }
  for (int i = 0; i < 3; ++i) p->ce_bin[i] = ce_bin ? ce_bin[i] : 1;
  if ((terms >> 6) & 1u) {
    if (!s_feat || !t_feat || feat_dim <= 0) return fail(CRDPN_E_BADARG, "crdpn_kd_mix: missing features");
  }
  p->s_feat = s_feat; p->t_feat = t_feat; p->feat_dim = (int)feat_dim;
  p->label = label; p->label_stride = (int)label_stride; p->n = (int)n; p->delta_bin = delta_bin; p->terms = terms;
  p->T = T; p->inv_T = 1.0f / T; p->w_kl = w_kl; p->w_rep = w_rep; p->w_gt = w_gt;
  p->ticket = (unsigned*)ws;
  p->row_loss = (float*)ws + 4;
  p->grad_loss = nullptr; p->loss_out = nullptr;
  for (int i = 0; i < 6; ++i) { p->d_s_out[i] = nullptr; p->d_t_out[i] = nullptr; }
  p->d_s_feat = nullptr; p->d_t_feat = nullptr;
  return CRDPN_OK;
}

extern "C" int crdpn_kd_mix_forward(const float* const* student_out, const float* const* teacher_out, const int32_t* widths,
                                    const float* student_feat, const float* teacher_feat, int64_t feat_dim,
                                    const float* label, int64_t label_stride, int64_t n, const int32_t* ce_bin,
                                    int32_t delta_bin, uint32_t terms, float temperature, float w_kl, float w_rep, float w_gt,
                                    float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  if (!loss) return fail(CRDPN_E_BADARG, "crdpn_kd_mix_forward: null loss");
  kdl::MixParams p;
  int rc = mix_fill(&p, student_out, teacher_out, widths, student_feat, teacher_feat, feat_dim, label, label_stride, n, ce_bin,
                    delta_bin, terms, temperature, w_kl, w_rep, w_gt, workspace, workspace_bytes);
  if (rc) return rc;
  p.loss_out = loss;
  cudaStream_t st = (cudaStream_t)stream;
  kdl::kd_mix_kernel<false><<<(unsigned)n, kdl::kMixThreads, 0, st>>>(p);
  CRDPN_LAUNCH_CHECK("kd_mix_kernel");
  return CRDPN_OK;
}

extern "C" int crdpn_kd_mix_backward(const float* const* student_out, const float* const* teacher_out, const int32_t* widths,
                                     const float* student_feat, const float* teacher_feat, int64_t feat_dim,
                                     const float* label, int64_t label_stride, int64_t n, const int32_t* ce_bin,
                                     int32_t delta_bin, uint32_t terms, float temperature, float w_kl, float w_rep, float w_gt,
                                     const float* grad_loss, float* const* d_student_out, float* const* d_teacher_out,
                                     float* d_student_feat, float* d_teacher_feat, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  kdl::MixParams p;
  int rc = mix_fill(&p, student_out, teacher_out, widths, student_feat, teacher_feat, feat_dim, label, label_stride, n, ce_bin,
                    delta_bin, terms, temperature, w_kl, w_rep, w_gt, workspace, workspace_bytes);
  if (rc) return rc;
  p.grad_loss = grad_loss;
  for (int i = 0; i < 6; ++i) {
    p.d_s_out[i] = d_student_out ? d_student_out[i] : nullptr;
    p.d_t_out[i] = d_teacher_out ? d_teacher_out[i] : nullptr;
  }
  p.d_s_feat = d_student_feat;
  p.d_t_feat = d_teacher_feat;
  kdl::kd_mix_kernel<true><<<(unsigned)n, kdl::kMixThreads, 0, (cudaStream_t)stream>>>(p);
  CRDPN_LAUNCH_CHECK("kd_mix_kernel");
  return CRDPN_OK;
}
