/*
 * crdpn_b200_dev.h -- DEVELOPMENT probes, built into libcrdpn_b200_dev.so (build.py), never into the product library
 * libcrdpn_b200.so.  Test infrastructure for hardware questions (tests/test_umma_tf32_probe_gpu.py); no reference
 * counterpart.
 */
#ifndef CRDPN_B200_DEV_H_
#define CRDPN_B200_DEV_H_
#include "crdpn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------------
 * Development probe (not a product path; no reference counterpart): one tile of the tensor-core formulation planned for the
 * bank-streaming CRD step -- tcgen05.mma kind::tf32 on an fp32 tile stored once in the K-major
 * SWIZZLE_128B image and read both K-major (scores = rows . [V2 | V1]^T) and MN-major (gradients^T = rows^T . C).
 * rows1, rows2 [64,128]; v1, v2 [48,128]; c1, c2 [64,48] (f32, device); out [128,192]: columns [0,96) scores of the 128
 * stacked rows (64 of bank 1, then 64 of bank 2) against [V2 | V1], [96,144) G2^T[e][b] = sum_r rows1[r][e] c2[r][b],
 * [144,192) G1^T[e][b] = sum_r rows2[r][e] c1[r][b].  mode 0: as described; mode 1: the score MMA with M = 64 (bank-1 rows only),
 * to record the TMEM placement of M = 64 accumulators (the dump still covers all 128 lanes); mode 2: both GEMMs with bf16
 * operands (kind::f16) from ONE K-major SWIZZLE_128B image, read K-major for the scores and MN-major for the gradients.
 * ------------------------------------------------------------------------------------------------- */
int crdpn_umma_tf32_probe(const float* rows1, const float* rows2, const float* v1, const float* v2, const float* c1,
                          const float* c2, float* out, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRDPN_B200_DEV_H_ */
