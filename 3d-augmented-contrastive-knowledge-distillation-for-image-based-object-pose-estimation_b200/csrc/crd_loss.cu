// crd_loss.cu -- the whole CRDLoss.forward / backward as ONE C call each.
//
// A CRD step on B200 is ~0.5 ms of device work made of a dozen launches; driven launch by launch from Python the
// host needs about as long again (tensor allocations, ctypes marshalling, context managers), and because the KD loop
// reads the loss back every step (KD/common/base_class.py:400-401) the device cannot run ahead of the host.  These two
// entry points enqueue every launch of the published CRDLoss.forward (embed heads -> negative draw -> fused
// score/loss/backward -> reduction + momentum update) and of its backward (the two embed-head backwards) back to back
// from C, so the host cost of a step is two foreign calls.
#include "common.cuh"

namespace crdpn {  // embed_kernels.cu
int embed_forward2(const float* xs, const float* Ws, const float* bs, int64_t s_dim, float* pre_s, float* v1, float* inv1,
                   const float* xt, const float* Wt, const float* bt, int64_t t_dim, float* pre_t, float* v2, float* inv2,
                   int64_t B, int64_t D, void* stream);
int embed_backward2(const float* xs, int64_t s_dim, const float* Ws, const float* v1, const float* inv1, const float* g1,
                    const float* xt, int64_t t_dim, const float* Wt, const float* v2, const float* inv2, const float* g2,
                    const float* scale, int64_t B, int64_t D, float* dWs, float* dbs, float* dxs, float* dWt, float* dbt,
                    float* dxt, float* d_pre, void* stream);
// crd_kernels.cu: scoring pass over the shard + reduction / momentum update / sum over ranks in one kernel
int sharded_step_core(void* bank1, void* bank2, int64_t row_stride, int bank_dtype, void* const* peer_bufs_host, int rank,
                      int world, int64_t Bmax, int64_t Dmax, const int64_t* contrast_idx, int64_t B, int64_t K1, int64_t D,
                      int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end, float T, float Z1, float Z2,
                      float eps, float momentum, float one_minus_momentum, float* v1_all, float* v2_all, int64_t* y_all,
                      float* partial, double* result, float* reduced, void* workspace, size_t workspace_bytes, int variant,
                      void* stream, int idx_mode = 0, uint64_t seed = 0, uint64_t offset = 0, int64_t draw_n = 0,
                      int64_t draw_base = 0, const float* v1_local = nullptr, const float* v2_local = nullptr,
                      const int64_t* y_local = nullptr, const int32_t* offs_host = nullptr);
extern thread_local const unsigned long long* g_sampler_offset_dev;   // crd_kernels.cu

// CUDA-graph replays: with variant bit 0x4000 (CRDPN_VARIANT_DEVICE_OFFSET) and the in-kernel uniform draw, `idx_scratch`
// is a device uint64 that is ADDED to `offset` by the kernels and advanced by B * K1 after the step, so that a captured
// step draws fresh negatives on every replay (the host-side `offset` is baked into the graph).
__global__ void sampler_offset_bump_kernel(unsigned long long* ctr, unsigned long long delta) { *ctr += delta; }
struct DevOffsetScope {
  explicit DevOffsetScope(const unsigned long long* p) { g_sampler_offset_dev = p; }
  ~DevOffsetScope() { g_sampler_offset_dev = nullptr; }
};
}  // namespace crdpn

using namespace crdpn;

extern "C" int crdpn_crd_loss_forward(
    const float* f_s, int64_t s_dim, const float* Ws, const float* bs,
    const float* f_t, int64_t t_dim, const float* Wt, const float* bt,
    const int64_t* y, const int64_t* contrast_idx,
    const float* alias_prob, const int64_t* alias_alias, uint64_t seed, uint64_t offset, int64_t* idx_scratch,
    void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
    int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
    float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
    float* pre_s, float* pre_t, float* v1, float* v2, float* inv1, float* inv2,
    double* result, float* grad_v1, float* grad_v2,
    void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!y) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward: null y");
  int rc = embed_forward2(f_s, Ws, bs, s_dim, pre_s, v1, inv1, f_t, Wt, bt, t_dim, pre_t, v2, inv2, B, D, stream);
  if (rc) return rc;
  const int64_t* idx = contrast_idx;
  if (idx == nullptr && alias_prob == nullptr && alias_alias == nullptr && !(variant & 0x200)) {
    // uniform sampler: the scoring pass draws the negatives itself (same Philox stream, same indices, no list in memory)
    unsigned long long* ctr = (variant & 0x4000) ? reinterpret_cast<unsigned long long*>(idx_scratch) : nullptr;
    if ((variant & 0x4000) && !ctr) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward: device sampler offset asked for, idx_scratch is NULL");
    DevOffsetScope scope(ctr);
    rc = crdpn_crd_step_drawn(bank1, bank2, row_stride, bank_dtype, v1, v2, y, B, K1, D, n_data, k_total, row_begin, row_end, T,
                              Z1, Z2, eps, momentum, one_minus_momentum, seed, offset, n_data, 0, result, grad_v1, grad_v2,
                              workspace, workspace_bytes, variant & ~0x4000, stream);
    if (rc == 0 && ctr) {
      sampler_offset_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctr, (unsigned long long)(B * K1));
      CRDPN_LAUNCH_CHECK("sampler_offset_bump_kernel");
    }
    return rc;
  }
  if (variant & 0x4000) return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_loss_forward: the device sampler offset needs the in-kernel uniform draw");
  if (idx == nullptr) {
    if (!idx_scratch) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward: contrast_idx is NULL and so is idx_scratch");
    rc = crdpn_alias_draw_contrast(alias_prob, alias_alias, n_data, y, B, K1, seed, offset, idx_scratch, stream);
    if (rc) return rc;
    idx = idx_scratch;
  }
  return crdpn_crd_step(bank1, bank2, row_stride, bank_dtype, v1, v2, idx, y, B, K1, D, n_data, k_total, row_begin,
                        row_end, T, Z1, Z2, eps, momentum, one_minus_momentum, result, grad_v1, grad_v2, workspace,
                        workspace_bytes, variant, stream);
}

extern "C" int crdpn_crd_loss_backward(
    const float* f_s, int64_t s_dim, const float* Ws, const float* v1, const float* inv1, const float* grad_v1,
    const float* f_t, int64_t t_dim, const float* Wt, const float* v2, const float* inv2, const float* grad_v2,
    const float* scale, int64_t B, int64_t D,
    float* dWs, float* dbs, float* dxs, float* dWt, float* dbt, float* dxt, float* d_pre_scratch, void* stream) {
  if (!d_pre_scratch) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_backward: null scratch");
  return embed_backward2(f_s, s_dim, Ws, v1, inv1, grad_v1, f_t, t_dim, Wt, v2, inv2, grad_v2, scale, B, D, dWs, dbs, dxs, dWt,
                         dbt, dxt, d_pre_scratch, stream);
}

// Row-sharded banks, one process per GPU: local embed heads -> all-gather of the anchors over NVLink peer memory ->
// (in-shard negative draw) -> scoring pass over this rank's shard + owner-only momentum update -> one-shot all-reduce of
// the packed partials [grad_v1 | grad_v2 | 8 result scalars], fused into the step's reduction kernel.  6 launches, one foreign call.
extern "C" int crdpn_crd_loss_forward_sharded(
    const float* f_s, int64_t s_dim, const float* Ws, const float* bs,
    const float* f_t, int64_t t_dim, const float* Wt, const float* bt,
    const int64_t* y_local, const int32_t* offs_host, void* const* peer_bufs_host, int rank, int world, int64_t Bmax,
    int64_t Dmax,
    const int64_t* contrast_idx, const float* alias_prob, const int64_t* alias_alias, uint64_t seed, uint64_t offset,
    int64_t* idx_scratch,
    void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
    int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
    float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
    float* pre_s, float* pre_t, float* v1_local, float* v2_local, float* inv1, float* inv2,
    float* v1_all, float* v2_all, int64_t* y_all, float* partial, double* result, float* reduced,
    void* workspace, size_t workspace_bytes, int variant, void* stream) {
  if (!y_local || !offs_host || !peer_bufs_host || !v1_all || !v2_all || !y_all || !partial || !result || !reduced)
    return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward_sharded: null pointer");
  if (world < 1 || rank < 0 || rank >= world) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward_sharded: bad rank / world");
  const int64_t B_loc = offs_host[rank + 1] - offs_host[rank], B = offs_host[world];
  if (B_loc <= 0 || B <= 0) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward_sharded: every rank must hold at least one anchor");
  int rc = embed_forward2(f_s, Ws, bs, s_dim, pre_s, v1_local, inv1, f_t, Wt, bt, t_dim, pre_t, v2_local, inv2, B_loc, D, stream);
  if (rc) return rc;
  const int64_t* idx = contrast_idx;
  const bool uniform_draw = idx == nullptr && alias_prob == nullptr && alias_alias == nullptr && !(variant & 0x200);
  if ((variant & 0x4000) && !(uniform_draw && idx_scratch))
    return fail(CRDPN_E_UNSUPPORTED, "crdpn_crd_loss_forward_sharded: the device sampler offset needs the in-kernel uniform draw");
  if (idx != nullptr || uniform_draw) {
    // the all-gather runs inside the core (with the filter pre-pass beside it); in-shard negatives, when asked for, are
    // drawn by the scoring pass / the pre-pass itself (uniform sampler)
    unsigned long long* ctr = (variant & 0x4000) ? reinterpret_cast<unsigned long long*>(idx_scratch) : nullptr;
    DevOffsetScope scope(ctr);
    variant &= ~0x4000;
    rc = sharded_step_core(bank1, bank2, row_stride, bank_dtype, peer_bufs_host, rank, world, Bmax, Dmax, idx, B, K1, D, n_data,
                             k_total, row_begin, row_end, T, Z1, Z2, eps, momentum, one_minus_momentum, v1_all, v2_all, y_all,
                             partial, result, reduced, workspace, workspace_bytes, variant & ~0x1000, stream,
                             uniform_draw ? 2 : ((variant & 0x1000) ? 1 : 0), seed, offset, row_end - row_begin, row_begin,
                             v1_local, v2_local, y_local, offs_host);
    if (rc == 0 && ctr) {
      sampler_offset_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctr, (unsigned long long)(B * K1));
      CRDPN_LAUNCH_CHECK("sampler_offset_bump_kernel");
    }
    return rc;
  }
  rc = crdpn_p2p_allgather_anchors(v1_local, v2_local, y_local, D, offs_host, peer_bufs_host, rank, world, Bmax, Dmax, v1_all,
                                   v2_all, y_all, stream);
  if (rc) return rc;
  if (idx == nullptr) {  // K1-1 negatives drawn inside this rank's shard; column 0 stays the global positive index
    if (!idx_scratch) return fail(CRDPN_E_BADARG, "crdpn_crd_loss_forward_sharded: contrast_idx is NULL and so is idx_scratch");
    rc = crdpn_alias_draw_contrast_local(alias_prob, alias_alias, row_end - row_begin, row_begin, y_all, B, K1, seed, offset,
                                         idx_scratch, stream);
    if (rc) return rc;
    idx = idx_scratch;
  }
  return sharded_step_core(bank1, bank2, row_stride, bank_dtype, peer_bufs_host, rank, world, Bmax, Dmax, idx, B, K1, D, n_data,
                           k_total, row_begin, row_end, T, Z1, Z2, eps, momentum, one_minus_momentum, v1_all, v2_all, y_all,
                           partial, result, reduced, workspace, workspace_bytes, variant & ~0x1000, stream,
                           (contrast_idx != nullptr && (variant & 0x1000)) ? 1 : 0);
}
