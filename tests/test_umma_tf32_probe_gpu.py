"""Hardware check of the TF32 tcgen05 building block planned for the bank-streaming CRD step (csrc/umma_tf32_probe.cu):
one fp32 tile in the K-major SWIZZLE_128B image, read K-major for the scores and MN-major for the gradients.
Tolerance: TF32 operands (10-bit mantissa), fp32 accumulation -> 2e-3 of each block's max."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_scores_and_gradients_from_one_tile(pkg, cuda):
    lib = pkg._native.dev_lib()
    rng = np.random.default_rng(46)
    rows1, rows2 = rng.normal(size=(64, 128)).astype(np.float32), rng.normal(size=(64, 128)).astype(np.float32)
    v1, v2 = rng.normal(size=(48, 128)).astype(np.float32), rng.normal(size=(48, 128)).astype(np.float32)
    c1, c2 = rng.normal(size=(64, 48)).astype(np.float32), rng.normal(size=(64, 48)).astype(np.float32)
    c1[rng.random(c1.shape) < 0.9] = 0     # the real coefficient matrices are sparse
    c2[rng.random(c2.shape) < 0.9] = 0
    d = [torch.from_numpy(a).to(cuda) for a in (rows1, rows2, v1, v2, c1, c2)]
    out = torch.full((128, 192), float("nan"), device=cuda)
    rc = lib.crdpn_umma_tf32_probe(*[t.data_ptr() for t in d], out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, pkg._native.lib().crdpn_last_error()
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    vcat = np.concatenate([v2, v1]).astype(np.float64)                      # [96, 128]
    want_s = np.concatenate([rows1, rows2]).astype(np.float64) @ vcat.T       # [128, 96]
    want_g2 = rows1.astype(np.float64).T @ c2.astype(np.float64)              # [128 features, 48]
    want_g1 = rows2.astype(np.float64).T @ c1.astype(np.float64)
    for name, g, w in (("scores", got[:, :96], want_s), ("G2^T", got[:, 96:144], want_g2), ("G1^T", got[:, 144:], want_g1)):
        err = np.abs(g - w).max() / np.abs(w).max()
        assert np.isfinite(g).all() and err < 2e-3, (name, err)


def test_m64_accumulator_placement(pkg, cuda):
    """M = 64 score MMA (the 64 bank-1 rows): which TMEM lanes hold which rows?  Recorded for the 32-row-tile variant."""
    lib = pkg._native.dev_lib()
    rng = np.random.default_rng(7)
    rows1, rows2 = rng.normal(size=(64, 128)).astype(np.float32), rng.normal(size=(64, 128)).astype(np.float32)
    v1, v2 = rng.normal(size=(48, 128)).astype(np.float32), rng.normal(size=(48, 128)).astype(np.float32)
    z = np.zeros((64, 48), np.float32)
    d = [torch.from_numpy(a).to(cuda) for a in (rows1, rows2, v1, v2, z, z)]
    out = torch.full((128, 192), float("nan"), device=cuda)
    assert lib.crdpn_umma_tf32_probe(*[t.data_ptr() for t in d], out.data_ptr(), 1, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)[:, :96]
    want = rows1.astype(np.float64) @ np.concatenate([v2, v1]).astype(np.float64).T     # [64, 96]
    scale = np.abs(want).max()
    lane_of_row = []
    for r in range(64):
        hits = [l for l in range(128) if np.abs(got[l] - want[r]).max() < 2e-3 * scale]
        assert len(hits) == 1, (r, hits)
        lane_of_row.append(hits[0])
    print("M=64 accumulator: row -> TMEM lane", lane_of_row)
    assert lane_of_row == list(range(64)) or lane_of_row == [16 * (r // 16) * 2 + r % 16 for r in range(64)], lane_of_row


def test_bf16_single_image_serves_both_gemms(pkg, cuda):
    """mode 2: 16-bit operands may be read MN-major from the ordinary SWIZZLE_128B image (no second copy of the tile)."""
    lib = pkg._native.dev_lib()
    rng = np.random.default_rng(3)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    rows1, rows2 = bf(rng.normal(size=(64, 128)).astype(np.float32)), bf(rng.normal(size=(64, 128)).astype(np.float32))
    v1, v2 = bf(rng.normal(size=(48, 128)).astype(np.float32)), bf(rng.normal(size=(48, 128)).astype(np.float32))
    c1, c2 = bf(rng.normal(size=(64, 48)).astype(np.float32)), bf(rng.normal(size=(64, 48)).astype(np.float32))
    d = [torch.from_numpy(a).to(cuda) for a in (rows1, rows2, v1, v2, c1, c2)]
    out = torch.full((128, 192), float("nan"), device=cuda)
    assert lib.crdpn_umma_tf32_probe(*[t.data_ptr() for t in d], out.data_ptr(), 2, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    vcat = np.concatenate([v2, v1]).astype(np.float64)
    want = {"scores": (got[:, :96], np.concatenate([rows1, rows2]).astype(np.float64) @ vcat.T),
            "G2^T": (got[:, 96:144], rows1.astype(np.float64).T @ c2.astype(np.float64)),
            "G1^T": (got[:, 144:], rows2.astype(np.float64).T @ c1.astype(np.float64))}
    for name, (g, w) in want.items():      # operands are exactly representable in bf16: only fp32 accumulation error
        err = np.abs(g - w).max() / np.abs(w).max()
        assert np.isfinite(g).all() and err < 1e-5, (name, err)
