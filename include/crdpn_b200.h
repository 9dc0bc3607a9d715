/*
 * crdpn_b200.h -- C ABI of libcrdpn_b200.so: the B200 (sm_100a) hot path of 3DAug-Pose's contrastive
 * distillation step: CRD memory-bank NCE (score + loss + closed-form backward + momentum update +
 * alias-method negative sampling) and the teacher's PointNet encoder (shared MLP + max-pool).
 *
 * The reference (/root/reference) is pure Python/PyTorch and has NO plugin / FFI interface; its only
 * extension seam is duck-typing on nn.Module.  Each entry point below therefore names
 * the reference interface (file:line) whose work it performs; the Python classes that call them mirror
 * those interfaces one-to-one (package crd.py / pointnet.py) and INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.  All *device* pointers unless marked host.
 *   - the caller owns every buffer (banks, workspaces, outputs); nothing is allocated or freed here.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, no hidden sync.
 *   - return 0 on success; CRDPN_E_* (>= 1000) for argument errors; otherwise the cudaError_t of the
 *     failed runtime call.  crdpn_last_error() returns a static description of the last failure of the
 *     calling thread.
 *   - bank shard: rows [row_begin,row_end) of the global [n_data, D] bank are resident; bank pointers
 *     address local row 0 == global row `row_begin`; contrast entries outside the shard are skipped.
 */
#ifndef CRDPN_B200_H_
#define CRDPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRDPN_ABI_VERSION 1

enum {
  CRDPN_OK = 0,
  CRDPN_E_BADARG = 1000,      /* null pointer / non-positive size */
  CRDPN_E_UNSUPPORTED = 1001, /* feature dimension or dtype without a compiled kernel */
  CRDPN_E_WORKSPACE = 1002,   /* workspace too small */
  CRDPN_E_ALIGN = 1003        /* pointer / stride not 16-byte aligned */
};

enum { CRDPN_F32 = 0, CRDPN_BF16 = 1 };

int crdpn_abi_version(void);
const char* crdpn_last_error(void);
/* number of kernels this library has launched since load (process-wide); feeds bench.py's gpu_launches */
uint64_t crdpn_launch_count(void);

/* Measurement aid (bench.py's roofline): when enabled, every launch of a dominant kernel (CRDPN_K_*) is
 * bracketed by cudaEventRecord on its own stream.  crdpn_timing_read synchronises those events, returns the
 * summed device time and launch count since the last read, and resets.  Off by default; no effect on results. */
enum { CRDPN_K_CRD_SCORE = 0, CRDPN_K_POINTNET_FWD = 1, CRDPN_K_COUNT = 2 };
int crdpn_timing_enable(int on);
int crdpn_timing_read(int kernel_id, double* total_ms, uint64_t* launches);

/* ---------------------------------------------------------------------------------------------------
 * Alias-method negative sampler.
 * Replaces: AliasMethod.__init__ / AliasMethod.draw of the published CRD algorithm (crd/memory.py in
 * HobbitLong/RepDistiller -- NOT vendored by the reference); in the reference the
 * consumer would be the KD loop at KD/common/base_class.py:339-387.
 * ------------------------------------------------------------------------------------------------- */
/* HOST pointers. Vose stack pairing in fp32; prob_out[n] fp32, alias_out[n] int64. */
int crdpn_alias_build(const float* probs_host, int64_t n, float* prob_out_host, int64_t* alias_out_host);
/* out[i] = draw(Philox4x32-10 block (seed, offset+i)), i < count.  prob == alias == NULL means the uniform tables
 * (prob[k] = 1 for all k, which is what uniform unigrams build): same indices, without the 4-byte gather per draw. */
int crdpn_alias_draw(const float* prob, const int64_t* alias, int64_t n, int64_t count,
                     uint64_t seed, uint64_t offset, int64_t* out, void* stream);
/* contrast_idx[B,K1]: as crdpn_alias_draw over B*K1 entries, then column 0 <- y[b]
 * (ContrastMemory.forward when idx is None). */
int crdpn_alias_draw_contrast(const float* prob, const int64_t* alias, int64_t n, const int64_t* y,
                              int64_t B, int64_t K1, uint64_t seed, uint64_t offset, int64_t* out,
                              void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Embed heads: flatten -> Linear(dim_in, D) -> x / ||x||_2  (published CRD `Embed` + `Normalize`, crd/criterion.py;
 * the reference's own heads are the `projector` MLPs, auxiliary/model.py:41-42, 238-241, whose 200-d output is what
 * these heads consume at KD/common/base_class.py:359,363).  x [B,dim_in], W [D,dim_in], b [D] f32.
 * forward : pre [B,D] = x W^T + b;  v [B,D] = pre / ||pre||_2 (no epsilon);  inv_norm [B] = 1 / ||pre||_2.
 * backward: given grad_v [B,D] and an optional device scalar `scale` (the upstream gradient of the loss; NULL = 1):
 *           dW [D,dim_in], db [D], dx [B,dim_in] (dx may be NULL: the teacher side needs none);
 *           d_pre [B,D] is caller-provided scratch.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_embed_forward(const float* x, const float* W, const float* b, int64_t B, int64_t dim_in, int64_t D,
                        float* pre, float* v, float* inv_norm, void* stream);
int crdpn_embed_backward(const float* x, const float* W, const float* v, const float* inv_norm, const float* grad_v,
                         const float* scale, int64_t B, int64_t dim_in, int64_t D, float* dW, float* db, float* dx,
                         float* d_pre, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * CRD scoring: fused gather . dot . exp . NCE loss . closed-form backward.  Never materialises the
 * B x K1 x D gathered tensor.
 * Replaces: ContrastMemory.forward (index_select + bmm + exp + div) and ContrastLoss.forward of the
 * published CRD algorithm and their autograd backward; reference insertion point
 * KD/vision/vanilla/vanilla_kd.py:158-160, called at KD/common/base_class.py:387.
 *
 *   s1[b,k] = <bank2[idx[b,k]], v1[b]>   s2[b,k] = <bank1[idx[b,k]], v2[b]>     e = exp(s/T)
 *   mode (Z1>0 && Z2>0): o = e/Z, loss_s/loss_t, grad_v1/grad_v2 (for upstream gradient 1)
 *   mode (otherwise)   : only sum(e1), sum(e2), count  (first call: Z = mean(e) * n_data)
 *
 * bank1/bank2: [row_end-row_begin, D] rows of `bank_dtype`, row_stride in ELEMENTS (both banks share it;
 *              an interleaved [N,2,D] allocation is bank2 = bank1 + D, row_stride = 2D).
 * v1,v2 [B,D] f32; contrast_idx [B,K1] int64 (column 0 = positive).  k_total = number of negatives per anchor
 * over ALL shards (the K of the NCE constant K*Pn); pass 0 for "K1-1" (single shard, or a replicated index list).
 * out_v1/out_v2: optional [B,K1] f32 (o, or raw e in sum mode; 0 for entries outside the shard).
 * result: 8 doubles {loss_s, loss_t, sum_e1, sum_e2, count, loss_s+loss_t, [float32 (loss_s+loss_t) in the first
 * 4 bytes of slot 6], 0}.  grad_v1/grad_v2 [B,D] f32 (written
 * only in full mode).  variant: 0 = default kernel; other values select tuning variants (bench only).
 * ------------------------------------------------------------------------------------------------- */
int crdpn_crd_workspace_bytes(int64_t B, int64_t K1, int64_t D, int device, size_t* bytes);
int crdpn_crd_score(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                    const float* v1, const float* v2, const int64_t* contrast_idx,
                    int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                    int64_t row_begin, int64_t row_end,
                    float T, float Z1, float Z2, float eps,
                    float* out_v1, float* out_v2, double* result, float* grad_v1, float* grad_v2,
                    void* workspace, size_t workspace_bytes, int variant, void* stream);

/* One full training step = crdpn_crd_score (full mode, no out_v) followed by crdpn_crd_momentum_update, in two
 * launches instead of three: the momentum update rides in the reduction launch (both only need the scoring
 * pass to have finished).  Z1, Z2 must already be frozen.  Same results, bit for bit, as the two calls. */
int crdpn_crd_step(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                   const float* v1, const float* v2, const int64_t* contrast_idx, const int64_t* y,
                   int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                   int64_t row_begin, int64_t row_end,
                   float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                   double* result, float* grad_v1, float* grad_v2,
                   void* workspace, size_t workspace_bytes, int variant, void* stream);

/* Where the contrast indices of a step may come from besides an int64 list:
 *   variant | 0x1000 (crdpn_crd_step, crdpn_crd_step_sharded): contrast_idx points at an INT32 [B,K1] list -- half the
 *     bytes to copy from the host and to scan (row indices of any bank that fits one GPU are < 2^31).
 *   crdpn_crd_step_drawn: no list at all -- the scoring pass draws entry (b, k) itself as
 *     draw_base + floor(u64(Philox4x32-10 block (seed, offset + b*K1 + k)) * draw_n / 2^64), column 0 = y[b]: bit for bit the
 *     list crdpn_alias_draw_contrast / _local writes for UNIFORM tables (prob = alias = NULL), i.e. the published
 *     ContrastMemory.forward(idx=None) with its all-ones unigram table, without 16 bytes of index traffic per entry and
 *     without the draw launch.  crdpn_crd_loss_forward(_sharded) take this route when contrast_idx, alias_prob and
 *     alias_alias are all NULL.  Not available for the bank-streaming kernels (variant | 0x200). */
int crdpn_crd_step_drawn(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                         const float* v1, const float* v2, const int64_t* y,
                         int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total,
                         int64_t row_begin, int64_t row_end,
                         float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                         uint64_t seed, uint64_t offset, int64_t draw_n, int64_t draw_base,
                         double* result, float* grad_v1, float* grad_v2,
                         void* workspace, size_t workspace_bytes, int variant, void* stream);

/* Band-sorted contrast lists (variant | 0x400 in crdpn_crd_step, crdpn_crd_step_drawn, crdpn_crd_step_sharded and the
 * crdpn_crd_loss_forward calls that end in them): a pre-pass (one CTA per anchor and 4096-entry chunk) keeps the entries whose
 * row this shard owns and STABLY sorts them by row band (32 bands of the shard); the scoring pass then gives the warps of an
 * anchor interleaved 32-entry blocks of those lists, so all warps sweep the bank together and a row that is drawn several times
 * per step comes from HBM once (headline shape: 2.91 -> 1.01 GB of DRAM traffic per step).  Same scores, so loss / gradients
 * equal the plain step's up to fp32 summation order, updated rows bit-identical, bit-reproducible run to run.  Silently falls
 * back to the plain step when the shape does not qualify (fewer than 4096 resident rows, more than 32 chunks per anchor,
 * more anchors than resident warps).  variant | 0x20 beside it promises that EVERY entry lives in this shard (in-shard
 * negatives): the pre-pass then skips its survivor compaction.  The Python mirror sets the bit by itself for shards of about
 * the L2's size and larger (ContrastMemory.sweep). */

/* Bank-STREAMING formulation of the same step (variant | 0x200 in crdpn_crd_step / crdpn_crd_loss_forward): the samples
 * are bucketed by bank tile and every resident tile passes through shared memory exactly once, so a row that is sampled
 * several times per step (B*K1 > resident rows) is read from HBM once.
 *   - bf16 banks: tcgen05 tensor-core kernel (csrc/crd_tc_stream.cuh; 64-row tiles by TMA, scores and gradients as bf16
 *     MMAs with fp32 accumulation in TMEM, integer slot counters -> order-independent results).  1.7x faster than the
 *     bf16 gather kernel at the headline shape; the Python mirror selects it automatically (ContrastMemory.streaming).
 *     Results within the bf16 tolerance (1e-2) of the gather formulation; updated bank rows bit-identical.
 *   - fp32 banks: EXPERIMENTAL register kernel (csrc/crd_stream.cuh; 32-row tiles), opt-in, slower than the gather
 *     kernel; same results to fp32 rounding (atomic scatter inside a tile: not bit-reproducible).
 * Requires D = 128, B <= 48, banks interleaved ([rows][2][D]: bank2 = bank1 + D, row_stride = 2D) or dense
 * (row_stride = D); otherwise CRDPN_E_UNSUPPORTED.  The workspace must hold crdpn_crd_stream_workspace_bytes. */
int crdpn_crd_stream_workspace_bytes(int64_t B, int64_t K1, int64_t D, int64_t rows_local, int device, size_t* bytes);

/* ---------------------------------------------------------------------------------------------------
 * The whole published CRDLoss.forward (crd/criterion.py: embed_s, embed_t, ContrastMemory.forward, the two
 * ContrastLoss terms) and its autograd backward, each as ONE call that enqueues every launch back to back:
 *   forward : both embed heads (2 launches) -> [crdpn_alias_draw_contrast when contrast_idx == NULL, into idx_scratch
 *             [B,K1] int64] -> crdpn_crd_step: 5 launches.  Same results, bit for bit, as the individual calls.
 *   backward: both embed-head backwards in 2 launches (dxs / dxt may be NULL; d_pre_scratch holds 2*B*D floats).
 * This is what the reference's KD loop would call at KD/common/base_class.py:387 (forward) and :394 (backward);
 * the host cost of a step drops from ~30 foreign calls / allocations to two.
 * variant | 0x4000 (CUDA-graph replays; in-kernel uniform draw only, i.e. contrast_idx = alias_prob = alias_alias = NULL):
 *   idx_scratch is then a DEVICE uint64 counter that the kernels add to `offset`, and a one-thread kernel advances it by
 *   B * K1 behind the step, so a captured step draws fresh negatives -- the ones the eager loop would draw -- on every replay.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_crd_loss_forward(
    const float* f_s, int64_t s_dim, const float* Ws, const float* bs,
    const float* f_t, int64_t t_dim, const float* Wt, const float* bt,
    const int64_t* y, const int64_t* contrast_idx,
    const float* alias_prob, const int64_t* alias_alias, uint64_t seed, uint64_t offset, int64_t* idx_scratch,
    void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
    int64_t B, int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
    float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
    float* pre_s, float* pre_t, float* v1, float* v2, float* inv1, float* inv2,
    double* result, float* grad_v1, float* grad_v2,
    void* workspace, size_t workspace_bytes, int variant, void* stream);
int crdpn_crd_loss_backward(
    const float* f_s, int64_t s_dim, const float* Ws, const float* v1, const float* inv1, const float* grad_v1,
    const float* f_t, int64_t t_dim, const float* Wt, const float* v2, const float* inv2, const float* grad_v2,
    const float* scale, int64_t B, int64_t D,
    float* dWs, float* dbs, float* dxs, float* dWt, float* dbt, float* dxt, float* d_pre_scratch, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Backward of the UNFUSED published surface ContrastMemory.forward -> (out_v1, out_v2) (crd/memory.py), for callers that
 * put their own criterion on the outputs (CRDLoss does not need it: crdpn_crd_step applies the closed-form NCE gradient in
 * the scoring pass).  With upstream gradients grad_out_v1/2 [B,K1] and the forward's outputs out_v1/2 [B,K1]:
 *   grad_v1[b] = sum_k grad_out_v1[b,k] out_v1[b,k] / T * bank2[idx[b,k]],  grad_v2[b] likewise with bank1.
 * The published code differentiates through a detached COPY of the gathered rows, i.e. the banks BEFORE the momentum
 * update of the same call: old_rows1/2 [B,D] f32 are the pre-update rows of y, substituted wherever a contrast index
 * equals some y[j].  Entries outside [row_begin,row_end) are skipped.  feat_dim % 4 == 0, <= 512; B <= 1024.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_crd_out_backward_workspace_bytes(int64_t B, int64_t K1, int64_t D, size_t* bytes);
int crdpn_crd_out_backward(const void* bank1, const void* bank2, int64_t row_stride, int bank_dtype,
                           const float* old_rows1, const float* old_rows2, const int64_t* y,
                           const int64_t* contrast_idx, const float* grad_out_v1, const float* grad_out_v2,
                           const float* out_v1, const float* out_v2, int64_t B, int64_t K1, int64_t D,
                           int64_t row_begin, int64_t row_end, float T, float* grad_v1, float* grad_v2,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Momentum update of both banks (ContrastMemory.forward, torch.no_grad block: index_select, mul_, add_,
 * pow/sum/pow, div, index_copy_).  bank[y] <- normalise(m*bank[y] + (1-m)*v), canonical reduction order
 * (oracle/crd_oracle.c), last duplicate of y wins, rows outside the shard untouched.
 * Must be enqueued after crdpn_crd_score on the same stream (scores read the pre-update banks).
 * ------------------------------------------------------------------------------------------------- */
int crdpn_crd_momentum_update(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                              const float* v1, const float* v2, const int64_t* y,
                              int64_t B, int64_t D, int64_t row_begin, int64_t row_end,
                              float momentum, float one_minus_momentum, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Exchanges of the row-sharded CRD step over NVLink peer memory (one process per GPU; no counterpart in the
 * single-GPU reference -- SURVEY.md section 8e).  Each rank allocates one exchange buffer, exports its CUDA IPC
 * handle (64 bytes, exchanged by the caller through any host channel), imports the peers' handles, and passes the
 * HOST array of the `world` base pointers (its own at index `rank`) to the two exchange kernels.  Bmax / Dmax fix the
 * buffer layout and must be the same on every rank.  world <= 8.  Every rank must issue the same sequence of calls
 * (any sequence: the payload areas are double-buffered by epoch parity, so a rank that runs ahead never overwrites words
 * a slower peer has not read).  A poll that lasts longer than CRDPN_P2P_TIMEOUT_S seconds (environment, default 600)
 * means a peer died or the sequences diverged: the kernel traps (sticky context error) instead of hanging the box.
 *   crdpn_p2p_allgather_anchors: local rows v1/v2 [b_loc,D] f32, y [b_loc] i64 -> all B rows in rank order;
 *                                offs_host[world+1] = prefix sums of the per-rank batch sizes.
 *   crdpn_p2p_allreduce_f32:     out[i] = sum over ranks (in rank order: same bits on every rank) of partial[i];
 *                                the payload is n_main floats followed by n_tail doubles (sent as floats: the step's
 *                                8 result scalars), out has n_main + n_tail floats.
 *   crdpn_p2p_allreduce_blocks:  IN-PLACE sum over ranks (rank order) of up to 4 blocks of float32 / float64 accumulators in
 *                                one launch (the hand-offs of the rank-synchronised PointNet training step,
 *                                crdpn_pointnet_sync_blocks); block_ptrs / counts / is_f64 are HOST arrays; the payload
 *                                (one LL word per float, two per double) must fit a reduction slot of the exchange buffer:
 *                                2*Bmax*Dmax + 8 + 8*Bmax words.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_p2p_allreduce_blocks(void* const* block_ptrs_host, const int64_t* counts_host, const int* is_f64_host, int n_blocks,
                               void* const* peer_bufs_host, int rank, int world, int64_t Bmax, int64_t Dmax, void* stream);
int crdpn_p2p_buffer_bytes(int64_t Bmax, int64_t Dmax, int world, size_t* bytes);
int crdpn_p2p_alloc(size_t bytes, void** dev_ptr);
int crdpn_p2p_free(void* dev_ptr);
int crdpn_p2p_export(void* dev_ptr, void* handle64_host);
int crdpn_p2p_import(const void* handle64_host, void** peer_ptr);
int crdpn_p2p_close(void* peer_ptr);
int crdpn_p2p_allgather_anchors(const float* v1, const float* v2, const int64_t* y, int64_t D,
                                const int32_t* offs_host, void* const* peer_bufs_host, int rank, int world,
                                int64_t Bmax, int64_t Dmax, float* out_v1, float* out_v2, int64_t* out_y, void* stream);
int crdpn_p2p_allreduce_f32(const float* partial, int64_t n_main, const double* tail_f64, int64_t n_tail, float* out,
                            void* const* peer_bufs_host, int rank, int world, int64_t Bmax, int64_t Dmax, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * In-batch contrastive KD losses: infoNCE_KD and poseNCE_KD,
 * auxiliary/model_utils.py:225-285, with rotation_err / angles_to_matrix, auxiliary/utils.py:156-202; call sites
 * training.py:57, KD/common/base_class.py:504-508.
 *   feat_pos <- dropout(feat_pos, p) (p = 0.3 in infoNCE_KD, 0 in poseNCE_KD); a = normalise(feat_ori), q = normalise(feat_pos)
 *   loss = mean_n -log( e^{a_n.q_n/tau} / (e^{a_n.q_n/tau} + sum_k w_nk e^{a_n.q_k/tau}) ), k over ALL rows including n
 *   weighting: 0 none (w = 1, infoNCE_KD); 1 linear, 2 square, 3 sqrt, 4 sin, 5 sinsin of rotation_err(label_n, label_k)/180
 *              (label [B,3] float32 degrees: azimuth, elevation, in-plane; w_nn = 0 exactly)
 * The dropout keep-mask is a counter-based stream: element e = n*C + c takes word e%4 of Philox4x32-10 block
 * (seed, offset + e/4); keep iff (word >> 8) * 2^-24 >= p; kept values are scaled by 1/(1-p).
 * Mode bits OR-ed into `weighting` select the file's other in-batch variants on the same kernels:
 *   0x100  negatives are the anchors' OWN rows, k = n excluded: sum_k becomes sum_{k != n} w_nk e^{a_n.a_k/tau}
 *          (infoNCE, model_utils.py:169-186, with weighting 0; poseNCE, :189-223, with a pose weighting)
 *   0x200  loss = mean_n -(a_n.q_n)/tau, the positive logit alone (singleinfoNCE_KD, :288-304)
 *   0x400  every k with k = n or rotation_err(label_n, label_k) <= 30 degrees is a positive:
 *          loss = mean_n -log( P_n / (P_n + sum_k e^{a_n.q_k/tau}) ), P_n = sum_{k positive} e^{a_n.q_k/tau} (multiposeNCE_KD, :307-351)
 * forward: 2 launches, writes the scalar loss; backward: 1 launch, gradients w.r.t. BOTH inputs (d_pos may be NULL),
 * scaled by the device scalar grad_loss (NULL = 1).  The workspace carries the normalised rows and soft weights from
 * forward to backward and must stay untouched in between.  Deterministic (fixed-order reductions).
 * ------------------------------------------------------------------------------------------------- */
int crdpn_nce_kd_workspace_bytes(int64_t B, int64_t C, size_t* bytes);
int crdpn_nce_kd_forward(const float* feat_ori, const float* feat_pos, const float* label, int64_t B, int64_t C,
                         float tau, int weighting, float dropout_p, uint64_t seed, uint64_t offset, float* loss,
                         void* workspace, size_t workspace_bytes, void* stream);
int crdpn_nce_kd_backward(const float* grad_loss, int64_t B, int64_t C, float tau, float dropout_p, uint64_t seed,
                          uint64_t offset, const void* workspace, size_t workspace_bytes, float* d_ori, float* d_pos,
                          void* stream);

/* ---------------------------------------------------------------------------------------------------
 * KD loss mixer: the scalar losses of one student / teacher step in ONE launch, and
 * their gradients in one more.  Replaces CELoss x3 + DeltaLoss (auxiliary/loss.py:7-34), TemperatureScaledKLDivLoss x7
 * (KD/vision/vanilla/vanilla_kd.py:8-32) and the weighted sum of calculate_kd_loss_new (vanilla_kd.py:143-164); call
 * site KD/common/base_class.py:365-387 (training.py:50-54 for the ground-truth part alone).
 *   loss = w_gt * [ sum_{i<3} CE(out_i, label_i // ce_bin_i) + SmoothL1(5*tanh(out_{3+i}[bin_i])/2, 5*((label_i % delta_bin)/delta_bin - .5)) ]
 *        + w_kl * sum_{i<6} T^2 KL(softmax(t_i/T) || softmax(s_i/T)) + w_rep * T^2 KL(softmax(tf/T) || softmax(sf/T))
 * (every term a batch mean, as in the reference).  `terms` selects the parts: bit i (0..5) KL of head i, bit 6 KL of the
 * features, bit 7+i (i<3) CE of head i, bit 10 the delta term; un-selected inputs may be NULL.  student_out /
 * teacher_out / d_* are HOST arrays of 6 device pointers ([n, widths[i]] f32 each), widths / ce_bin HOST arrays.
 * label: [n, label_stride] f32 degrees (column i = angle i).  Gradient outputs may individually be NULL.
 * workspace: crdpn_kd_mix_workspace_bytes(n) bytes whose first 16 bytes are ZERO before the first call (a ticket that
 * every call leaves zeroed).
 * ------------------------------------------------------------------------------------------------- */
int crdpn_kd_mix_workspace_bytes(int64_t n, size_t* bytes);
int crdpn_kd_mix_forward(const float* const* student_out, const float* const* teacher_out, const int32_t* widths,
                         const float* student_feat, const float* teacher_feat, int64_t feat_dim,
                         const float* label, int64_t label_stride, int64_t n, const int32_t* ce_bin,
                         int32_t delta_bin, uint32_t terms, float temperature, float w_kl, float w_rep, float w_gt,
                         float* loss, void* workspace, size_t workspace_bytes, void* stream);
int crdpn_kd_mix_backward(const float* const* student_out, const float* const* teacher_out, const int32_t* widths,
                          const float* student_feat, const float* teacher_feat, int64_t feat_dim,
                          const float* label, int64_t label_stride, int64_t n, const int32_t* ce_bin,
                          int32_t delta_bin, uint32_t terms, float temperature, float w_kl, float w_rep, float w_gt,
                          const float* grad_loss, float* const* d_student_out, float* const* d_teacher_out,
                          float* d_student_feat, float* d_teacher_feat, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Point-cloud input producer: read_pointcloud, auxiliary/dataset.py:121-150 (call sites
 * dataset.py:299,607), for a whole batch in one launch.  The raw mesh vertices of all M models are resident:
 * vertices [sum_m V_m, 3] float64 (as pymesh yields them), cloud_offsets [M+1] (prefix sums of V_m, device).
 * For b < B: cloud cid = cloud_ids[b]; P of its V vertices are taken without replacement -- rows subset[b, :] when
 * `subset` is given, else the first P images of the keyed Feistel permutation of [0, V) (key: Philox4x32-10 blocks
 * 2s and 2s+1 of `seed`, s = offset + b; the chosen rows are written to subset_out when non-NULL) --, rotated about z by
 * rotation_deg[b] degrees when that is non-zero (float64, point_cloud @ R^T), cast to float32, transposed to [3, P],
 * shifted by the global minimum and divided by the global maximum of the shifted cloud.  out [B, 3, P] float32 in [0, 1].
 * A cloud with fewer than P vertices (the reference raises) is filled with NaN.  P <= 8192.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_pointcloud_sample(const double* vertices, const int64_t* cloud_offsets, const int64_t* cloud_ids,
                            const float* rotation_deg, const int64_t* subset, uint64_t seed, uint64_t offset,
                            int64_t B, int64_t P, float* out, int64_t* subset_out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * PoseEstimator tail: cat(shape, img) -> DeformNet (2048 -> 2048 -> 1024 -> 512 -> 200,
 * three BatchNorm1d + ReLU, tanh) -> six fc_* heads, and projector(img) (1024 -> 800 -> 400 -> 200, two BatchNorm1d).
 * Replaces: /root/reference/auxiliary/model.py:183-203 (DeformNet.forward), :260-272 (the tail of PoseEstimator.forward);
 * eval-mode call KD/common/base_class.py:363, train-mode call training.py:47 (model.train() at :30).
 *
 * The whole chain is ONE launch of a persistent tcgen05 kernel (csrc/pose_tail.cu): every layer is a set of
 * (128-output tile, K-split) tasks over all SMs, layers hand over through device-side counters, weights stream once from
 * HBM as bf16 (hi, lo) operand images -- three MMAs per product, fp32-accurate (outputs within 1e-5 of the fp32
 * reference); CRDPN_POSE_TAIL_BF16 reads the hi planes only (half the bytes; north_star's 1e-2 tolerance mode).
 * A chain is described by up to 10 layers; layer l computes act(BN(x_src W^T + bias)):
 *   weights  : image made by crdpn_pose_tail_pack_weights from W [O, I] fp32 (row_scale: eval-mode BatchNorm folded in)
 *   src      : -1 = cat(shape_feature, img_feature), -2 = img_feature, k >= 0 = the output of layer k (k < l)
 *   act      : 0 none, 1 ReLU, 2 tanh
 *   out      : [B, O] fp32 or NULL (intermediate activations are kept as operand images in the workspace only)
 *   gamma .. : with CRDPN_POSE_TAIL_TRAIN and gamma != NULL the layer applies batch-statistics BatchNorm (biased variance
 *              over the B rows, eps), writes save_mean / save_istd [O] and xhat [B, O] (may be NULL) for the backward, and
 *              updates running_mean / running_var in place (momentum, unbiased variance) when they are non-NULL.
 * B <= 256 rows per call.  The workspace (crdpn_pose_tail_workspace_bytes, 1024-byte aligned) must be ZERO when first used
 * (its first 1 KB holds the hand-over counters, which the kernel leaves zero again) and belongs to one stream at a time.
 * ------------------------------------------------------------------------------------------------- */
typedef struct crdpn_pose_tail_layer {
  const void* weights;
  const float* bias;
  int64_t O, I;
  int32_t src, act;
  float* out;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* save_mean;
  float* save_istd;
  float* xhat;
} crdpn_pose_tail_layer;
/* CRDPN_POSE_TAIL_PROF (debugging aid): per task, eight %globaltimer stamps (accumulator ready, TMEM drained, partial stored, tile
 * complete, reduce loop left, after the memory fence, after the proxy fence, task done) as uint64 [n_tasks][8] at byte 1024 of the workspace; tasks are numbered in dependency order. */
enum { CRDPN_POSE_TAIL_BF16 = 1, CRDPN_POSE_TAIL_TRAIN = 2, CRDPN_POSE_TAIL_PROF = 4 };
int crdpn_pose_tail_image_bytes(int64_t O, int64_t I, size_t* bytes);
int crdpn_pose_tail_pack_weights(const float* W, const float* row_scale, int64_t O, int64_t I, void* image, void* stream);
/* layers: HOST array */
int crdpn_pose_tail_workspace_bytes(const crdpn_pose_tail_layer* layers, int n_layers, int64_t B, int64_t shape_dim,
                                    int64_t img_dim, size_t* bytes);
int crdpn_pose_tail_forward(const crdpn_pose_tail_layer* layers, int n_layers, const float* shape_feature,
                            const float* img_feature, int64_t B, int64_t shape_dim, int64_t img_dim, int flags,
                            float bn_momentum, float bn_eps, void* workspace, size_t workspace_bytes, void* stream);

/* Train-mode backward of the same chain (what autograd does for training.py:75) as ONE call: per layer, last to first, the
 * activation / batch-statistics BatchNorm pull-back (d_gamma, d_beta, d_bias ride in it), dW = g^T x_in, and dx = g W
 * accumulated into the gradient of the layer's source; fp32 FFMA kernels, 3-7 launches per layer (the dx products run split-K
 * with a fixed-order sum of the partials: deterministic).
 *   y     : the layer's forward output [B, O] (crdpn_pose_tail_layer.out of the forward call)
 *   xhat, istd, gamma : the forward's saved BatchNorm values (NULL gamma: the layer has no BatchNorm)
 *   g_out : gradient w.r.t. this layer's output coming from OUTSIDE the chain (the losses), or NULL
 *   W     : the raw fp32 weights [O, I];  dW [O, I], db [O], dgamma / dbeta [O] are written (not accumulated)
 * d_shape_feature [B, shape_dim] / d_img_feature [B, img_dim] may be NULL.  Workspace: crdpn_pose_tail_backward_workspace_bytes. */
typedef struct crdpn_pose_tail_bwd_layer {
  const float* W;
  const float* y;
  const float* xhat;
  const float* gamma;
  const float* istd;
  const float* g_out;
  float* dW;
  float* db;
  float* dgamma;
  float* dbeta;
  int64_t O, I;
  int32_t src, act;
} crdpn_pose_tail_bwd_layer;
int crdpn_pose_tail_backward_workspace_bytes(const crdpn_pose_tail_bwd_layer* layers, int n_layers, int64_t B, size_t* bytes);
int crdpn_pose_tail_backward(const crdpn_pose_tail_bwd_layer* layers, int n_layers, const float* shape_feature,
                             const float* img_feature, int64_t B, int64_t shape_dim, int64_t img_dim,
                             float* d_shape_feature, float* d_img_feature, void* workspace, size_t workspace_bytes, void* stream);

/* The sharded step's forward as ONE call (one process per GPU, exchanges over NVLink peer memory as above): both embed
 * heads on the LOCAL anchors -> crdpn_p2p_allgather_anchors -> [contrast_idx == NULL: crdpn_alias_draw_contrast_local,
 * K1-1 negatives per anchor inside this rank's shard] -> crdpn_crd_step over the shard (gradients into `partial`
 * [2*B*D]) whose reduction kernel also sums over the ranks into `reduced` [2*B*D + 8] (the 8 result scalars ride behind
 * the gradients as fp32 words; word 5 is the loss of the whole batch).  6 launches.  The backward is crdpn_crd_loss_backward on the local
 * rows of `reduced`.  offs_host[world+1] = prefix sums of the per-rank batch sizes; B = offs_host[world]. */
int crdpn_alias_draw_contrast_local(const float* prob, const int64_t* alias, int64_t n_local, int64_t row_base,
                                    const int64_t* y, int64_t B, int64_t K1, uint64_t seed, uint64_t offset,
                                    int64_t* out, void* stream);
int crdpn_crd_loss_forward_sharded(
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
    void* workspace, size_t workspace_bytes, int variant, void* stream);

/* The sharded step on device-resident embeddings (the per-rank unit of BASELINE configs[3]: K negatives against an
 * N-row bank sharded over `world` GPUs): crdpn_p2p_allgather_anchors of the local rows -> scoring pass over this rank's
 * rows [row_begin,row_end) of the replicated (or per-shard) contrast_idx [B,K1] -> ONE kernel that reduces the per-warp
 * partials, momentum-updates the positive rows this rank owns and sums gradients / loss partials over the ranks (LL
 * words pushed into every peer's slot and summed in rank order: identical bits on every rank).  3 launches, no NCCL.
 * Outputs: v1_all / v2_all [B,D], y_all [B]; reduced [2*B*D + 8] f32 = grad_v1 | grad_v2 | {loss_s, loss_t, -, -, -,
 * loss_s + loss_t, -, -} of the WHOLE batch over ALL shards; partial [2*B*D] and result [8] are this rank's shares.
 * With variant | 0x200 (bank-streaming kernels) the sum over ranks is a separate crdpn_p2p_allreduce_f32 launch. */
int crdpn_crd_step_sharded(void* bank1, void* bank2, int64_t row_stride, int bank_dtype,
                           const float* v1_local, const float* v2_local, const int64_t* y_local,
                           const int32_t* offs_host, void* const* peer_bufs_host, int rank, int world,
                           int64_t Bmax, int64_t Dmax, const int64_t* contrast_idx,
                           int64_t K1, int64_t D, int64_t n_data, int64_t k_total, int64_t row_begin, int64_t row_end,
                           float T, float Z1, float Z2, float eps, float momentum, float one_minus_momentum,
                           float* v1_all, float* v2_all, int64_t* y_all, float* partial, double* result, float* reduced,
                           void* workspace, size_t workspace_bytes, int variant, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * PointNet encoder, eval-mode BatchNorm (the KD-time teacher, KD/common/base_class.py:317,363).
 * Replaces: ShapeEncoderPC.forward, auxiliary/model.py:174-180 (conv1/bn1/relu, conv2/bn2/relu,
 * conv3/bn3, max over points), BN folded into the weights by crdpn_pointnet_pack.
 * ------------------------------------------------------------------------------------------------- */
/* Fold BN (eval statistics) into the three 1x1 convolutions and pack the bf16 tensor-core operand images.
 * All inputs f32 device pointers with the reference's state_dict shapes (model.py:162-172);
 * F = feature_dim in {128,256,512,1024}. packed: crdpn_pointnet_packed_bytes(F) bytes, 16-byte aligned. */
int crdpn_pointnet_packed_bytes(int64_t F, size_t* bytes);
int crdpn_pointnet_pack(const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                        const float* conv3_w, const float* conv3_b,
                        const float* bn1_w, const float* bn1_b, const float* bn1_mean, const float* bn1_var,
                        const float* bn2_w, const float* bn2_b, const float* bn2_mean, const float* bn2_var,
                        const float* bn3_w, const float* bn3_b, const float* bn3_mean, const float* bn3_var,
                        float bn_eps, int64_t F, void* packed, void* stream);
/* x [B,3,P] f32 channel-first contiguous; out [B,F] f32. workspace: crdpn_pointnet_workspace_bytes. */
int crdpn_pointnet_workspace_bytes(int64_t B, int64_t P, int64_t F, int device, size_t* bytes);
int crdpn_pointnet_forward_eval(const float* x, int64_t B, int64_t P, int64_t F, const void* packed,
                                float* out, void* workspace, size_t workspace_bytes, int variant,
                                void* stream);

/* ---------------------------------------------------------------------------------------------------
 * PointNet encoder, train-mode BatchNorm (teacher training: model.train() at training.py:30, forward at
 * training.py:47 -> auxiliary/model.py:257 -> ShapeEncoderPC.forward model.py:174-180; backward via
 * loss.backward() at training.py:75).
 * BatchNorm uses the statistics of this batch (over all B*P points, biased variance); running_mean / running_var
 * are updated in place with `bn_momentum` (unbiased variance) and num_batches_tracked += 1, exactly as
 * nn.BatchNorm1d does.  `ctx` is a caller-owned buffer (crdpn_pointnet_train_ctx_bytes, 1024-byte aligned) that
 * carries what backward needs (batch statistics, arg-max point per (cloud, channel), h2 in bf16); it must stay
 * untouched between the forward and its backward.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_pointnet_train_ctx_bytes(int64_t B, int64_t P, int64_t F, size_t* bytes);
int crdpn_pointnet_forward_train(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
    const float* conv3_w, const float* conv3_b,
    const float* bn1_w, const float* bn1_b, float* bn1_mean, float* bn1_var, int64_t* bn1_num_batches_tracked,
    const float* bn2_w, const float* bn2_b, float* bn2_mean, float* bn2_var, int64_t* bn2_num_batches_tracked,
    const float* bn3_w, const float* bn3_b, float* bn3_mean, float* bn3_var, int64_t* bn3_num_batches_tracked,
    float bn_eps, float bn_momentum, float* out, void* ctx, size_t ctx_bytes, int variant, void* stream);
/* Backward of the train-mode forward (autograd of model.py:174-180): grad_out [B,F] f32 -> gradients of the 12
 * parameter tensors (same shapes as the parameters; the input x needs no gradient -- it is data).  The max over
 * points makes layer 3's backward sparse (one point per (cloud, channel)) plus a rank-structured dense term from
 * BatchNorm's batch coupling, so the B*F*P tensor is never formed here either.
 * workspace: crdpn_pointnet_backward_workspace_bytes, 1024-byte aligned, scratch. */
int crdpn_pointnet_backward_workspace_bytes(int64_t B, int64_t P, int64_t F, size_t* bytes);
int crdpn_pointnet_backward(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv2_w, const float* conv3_w,
    const float* bn1_w, const float* bn1_b, const float* bn2_w, const float* bn2_b,
    const float* bn3_w, const float* bn3_b,
    const float* grad_out, const void* ctx, size_t ctx_bytes,
    float* d_conv1_w, float* d_conv1_b, float* d_conv2_w, float* d_conv2_b, float* d_conv3_w, float* d_conv3_b,
    float* d_bn1_w, float* d_bn1_b, float* d_bn2_w, float* d_bn2_b, float* d_bn3_w, float* d_bn3_b,
    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Rank-synchronised (SyncBN-style) PointNet training over several GPUs (no counterpart in the
 * single-GPU reference).  The *_phased variants run phases [phase_begin, phase_end) of the same launch sequence as
 * crdpn_pointnet_forward_train / crdpn_pointnet_backward (which are phases 0..4 with total_points = B*P) and take the
 * number of points of ALL ranks, over which the batch statistics are defined.  Between phase k and k+1 (k = 0, 1, 2) the
 * caller SUMS over ranks the accumulators that crdpn_pointnet_sync_blocks(sync_point = k for the forward, 3 + k for the
 * backward) names: buffer 0 = train ctx, 1 = backward workspace, 2 = d_bn3_w, 3 = d_bn3_b; byte offset, element count and
 * float64 flag per block (at most 4 blocks).  Every rank then holds the statistics of the global batch, and the parameter
 * gradients come out already summed over ranks (identical on every rank): no separate gradient all-reduce.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_pointnet_sync_blocks(int64_t B, int64_t P, int64_t F, int sync_point, int* n_blocks, int* buffer,
                               size_t* byte_offset, int64_t* count, int* is_f64);
int crdpn_pointnet_forward_train_phased(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
    const float* conv3_w, const float* conv3_b,
    const float* bn1_w, const float* bn1_b, float* bn1_mean, float* bn1_var, int64_t* bn1_num_batches_tracked,
    const float* bn2_w, const float* bn2_b, float* bn2_mean, float* bn2_var, int64_t* bn2_num_batches_tracked,
    const float* bn3_w, const float* bn3_b, float* bn3_mean, float* bn3_var, int64_t* bn3_num_batches_tracked,
    float bn_eps, float bn_momentum, float* out, void* ctx, size_t ctx_bytes, int variant,
    int phase_begin, int phase_end, int64_t total_points, void* stream);
int crdpn_pointnet_backward_phased(
    const float* x, int64_t B, int64_t P, int64_t F,
    const float* conv1_w, const float* conv2_w, const float* conv3_w,
    const float* bn1_w, const float* bn1_b, const float* bn2_w, const float* bn2_b,
    const float* bn3_w, const float* bn3_b,
    const float* grad_out, const void* ctx, size_t ctx_bytes,
    float* d_conv1_w, float* d_conv1_b, float* d_conv2_w, float* d_conv2_b, float* d_conv3_w, float* d_conv3_b,
    float* d_bn1_w, float* d_bn1_b, float* d_bn2_w, float* d_bn2_b, float* d_bn3_w, float* d_bn3_b,
    void* workspace, size_t workspace_bytes, int phase_begin, int phase_end, int64_t total_points, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRDPN_B200_H_ */
