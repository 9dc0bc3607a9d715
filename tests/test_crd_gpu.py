"""GPU parity tests of the CRD hot path: CUDA kernels (through the C ABI / the CRDLoss module) vs the CPU oracle.

Tolerances (BASELINE.json north_star): indices and memory-row updates bit-exact; loss, gradients and
scores within 1e-4 relative in fp32, 1e-2 with bf16 banks.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL32 = 1e-4
RELBF = 1e-2


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / (np.abs(want).max() + 1e-30)


def _case(B, K1, D, N, seed=0, dup_positive=False):
    g = torch.Generator().manual_seed(seed)
    stdv = 1.0 / (D / 3) ** 0.5
    bank = (torch.rand(N, 2, D, generator=g) * 2 * stdv - stdv)
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1)
    y = torch.randperm(N, generator=g)[:B] if B <= N else torch.randint(0, N, (B,), generator=g)
    idx = torch.randint(0, N, (B, K1), generator=g)
    idx[:, 0] = y
    if dup_positive and K1 > 3 and B > 1:
        idx[1, 3] = y[0]  # anchor 0's positive row is anchor 1's negative
        idx[0, 2] = idx[0, 1]  # repeated negative
    return bank, v1, v2, y, idx


def _score(pkg, dev, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=0, row_end=None, want_out=True, variant=0,
           dtype=torch.float32, interleaved=True, k_total=0):
    """Direct C-ABI call. `bank` is the FULL [N,2,D] fp32 tensor; a CPU bank's shard slice is uploaded, a bank that
    already lives on the device (fp32, interleaved) is used in place."""
    lib = pkg._native.lib()
    row_end = N if row_end is None else row_end
    B, K1 = idx.shape
    D = v1.shape[1]
    shard = bank[row_begin:row_end].to(dtype)
    if bank.is_cuda:
        assert interleaved and dtype == torch.float32
        b1, b2, stride = shard[:, 0, :], shard[:, 1, :], 2 * D
    elif interleaved:
        dbank = shard.contiguous().to(dev)
        b1, b2, stride = dbank[:, 0, :], dbank[:, 1, :], 2 * D
    else:
        b1, b2, stride = shard[:, 0, :].contiguous().to(dev), shard[:, 1, :].contiguous().to(dev), D
    dv1, dv2, didx = v1.to(dev), v2.to(dev), idx.to(dev)
    n = ctypes.c_size_t(0)
    pkg._native.check(lib.crdpn_crd_workspace_bytes(B, K1, D, 0, ctypes.byref(n)), "ws")
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
    res = torch.full((8,), -7.0, dtype=torch.float64, device=dev)
    g1 = torch.full((B, D), float("nan"), device=dev)
    g2 = torch.full((B, D), float("nan"), device=dev)
    o1 = torch.full((B, K1), float("nan"), device=dev) if want_out else None
    o2 = torch.full((B, K1), float("nan"), device=dev) if want_out else None
    rc = lib.crdpn_crd_score(b1.data_ptr() if row_end > row_begin else None, b2.data_ptr() if row_end > row_begin else None,
                             stride, 0 if dtype == torch.float32 else 1,
                             dv1.data_ptr(), dv2.data_ptr(), didx.data_ptr(), B, K1, D, N, k_total, row_begin, row_end,
                             T, Z1, Z2, 1e-7, o1.data_ptr() if want_out else None, o2.data_ptr() if want_out else None,
                             res.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), ws.numel(), variant,
                             torch.cuda.current_stream().cuda_stream)
    pkg._native.check(rc, "crdpn_crd_score")
    torch.cuda.synchronize()
    return dict(res=res.cpu().numpy(), g1=g1.cpu().numpy(), g2=g2.cpu().numpy(),
                o1=None if o1 is None else o1.cpu().numpy(), o2=None if o2 is None else o2.cpu().numpy())


def _oracle_score(oracle, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=0, row_end=None, dtype=torch.float32):
    row_end = N if row_end is None else row_end
    sh = bank[row_begin:row_end].to(dtype).float().numpy()
    return oracle.crd_score(sh[:, 0, :], sh[:, 1, :], v1.numpy(), v2.numpy(), idx.numpy(), N, T, Z1, Z2,
                            row_begin=row_begin, row_end=row_end)


def _check_full(got, want, rel):
    # loss: relative, with a 1e-6 absolute floor (with no negatives the loss is ~eps/o ~ 1e-7, i.e. pure
    # fp32 rounding of log(o/(o+eps)) in any fp32 implementation, the stock one included)
    assert abs(got["res"][0] - want["loss_s"]) < rel * (abs(want["loss_s"]) + 1e-2)
    assert abs(got["res"][1] - want["loss_t"]) < rel * (abs(want["loss_t"]) + 1e-2)
    assert _rel(got["res"][2], want["sum_e1"]) < rel and _rel(got["res"][3], want["sum_e2"]) < rel
    assert got["res"][4] == want["count"]
    assert _rel(got["g1"], want["grad_v1"]) < rel and _rel(got["g2"], want["grad_v2"]) < rel
    if got["o1"] is not None:
        assert _rel(got["o1"], want["out_v1"]) < rel and _rel(got["o2"], want["out_v2"]) < rel
        assert np.array_equal(got["o1"] == 0, want["out_v1"] == 0)  # same set of skipped (out-of-shard) entries


@pytest.mark.parametrize("B,K1,D,N", [
    (1, 1, 128, 16),          # a single positive, no negatives
    (3, 2, 128, 8),           # fewer pairs than warps
    (5, 33, 64, 100),         # ragged: K1 not a multiple of 32
    (7, 257, 128, 1000),
    (46, 1025, 128, 5000),    # many anchors per warp boundary
    (4, 4097, 256, 3000),
    (9, 130, 32, 64),
    (2, 70, 512, 300),
    (300, 17, 128, 500),      # more anchors than CTAs per anchor: several anchors per warp range
    (6000, 3, 64, 100),       # each warp range spans several anchors (maxseg > 2); duplicate positives
])
def test_score_parity_shapes(pkg, oracle, cuda, B, K1, D, N):
    bank, v1, v2, y, idx = _case(B, K1, D, N, seed=B + K1, dup_positive=True)
    T, Z1, Z2 = 0.07, 37.5 * N / 100, 21.0 * N / 100
    want = _oracle_score(oracle, bank, v1, v2, idx, N, T, Z1, Z2)
    got = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2)
    _check_full(got, want, REL32)
    # separately allocated (non-interleaved) banks give bit-identical results
    got2 = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, interleaved=False)
    for k in ("res", "g1", "g2", "o1", "o2"):
        assert np.array_equal(got[k], got2[k]), k


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7])
def test_score_parity_all_tuning_variants_d128(pkg, oracle, cuda, variant):
    bank, v1, v2, y, idx = _case(11, 777, 128, 2000, seed=variant, dup_positive=True)
    T, Z1, Z2 = 0.07, 600.0, 900.0
    want = _oracle_score(oracle, bank, v1, v2, idx, 2000, T, Z1, Z2)
    got = _score(pkg, cuda, bank, v1, v2, idx, 2000, T, Z1, Z2, variant=variant)
    _check_full(got, want, REL32)


@pytest.mark.parametrize("D,variant", [(64, 2), (256, 2)])
def test_score_parity_other_variants(pkg, oracle, cuda, D, variant):
    bank, v1, v2, y, idx = _case(6, 300, D, 700, seed=D)
    want = _oracle_score(oracle, bank, v1, v2, idx, 700, 0.07, 300.0, 200.0)
    got = _score(pkg, cuda, bank, v1, v2, idx, 700, 0.07, 300.0, 200.0, variant=variant)
    _check_full(got, want, REL32)


def test_sum_mode_first_call(pkg, oracle, cuda):
    bank, v1, v2, y, idx = _case(8, 513, 128, 4000, seed=3)
    want = _oracle_score(oracle, bank, v1, v2, idx, 4000, 0.07, -1.0, -1.0)
    got = _score(pkg, cuda, bank, v1, v2, idx, 4000, 0.07, -1.0, -1.0)
    assert got["res"][0] == 0 and got["res"][1] == 0
    assert _rel(got["res"][2], want["sum_e1"]) < REL32 and _rel(got["res"][3], want["sum_e2"]) < REL32
    assert got["res"][4] == 8 * 513
    assert _rel(got["o1"], want["out_v1"]) < REL32   # raw exponentials in sum mode
    assert np.isnan(got["g1"]).all()                  # gradients are not touched in sum mode


def test_sharded_partials_sum_to_unsharded_and_match_oracle(pkg, oracle, cuda):
    N, R = 1003, 4
    bank, v1, v2, y, idx = _case(10, 400, 128, N, seed=9, dup_positive=True)
    T, Z1, Z2 = 0.07, 300.0, 350.0
    full = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2)
    acc = None
    for r in range(R):
        lo, hi = N * r // R, N * (r + 1) // R
        part = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=lo, row_end=hi)
        want = _oracle_score(oracle, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=lo, row_end=hi)
        _check_full(part, want, REL32)
        acc = part if acc is None else {k: acc[k] + part[k] for k in part}
    for k in ("res", "g1", "g2", "o1", "o2"):
        assert _rel(acc[k], full[k]) < 1e-5, k
    # an empty shard contributes exactly nothing
    empty = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=500, row_end=500)
    assert not empty["res"][:5].any() and not empty["g1"].any() and not empty["o1"].any()


def test_deterministic_bitwise(pkg, cuda):
    bank, v1, v2, y, idx = _case(46, 2049, 128, 9000, seed=1)
    a = _score(pkg, cuda, bank, v1, v2, idx, 9000, 0.07, 2000.0, 2500.0)
    b = _score(pkg, cuda, bank, v1, v2, idx, 9000, 0.07, 2000.0, 2500.0)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_bf16_banks_within_bf16_tolerance(pkg, oracle, cuda):
    bank, v1, v2, y, idx = _case(12, 600, 128, 3000, seed=4)
    T, Z1, Z2 = 0.07, 1000.0, 1200.0
    # kernel on bf16 banks vs the oracle on the same bf16-rounded rows: fp32-level agreement
    want_same = _oracle_score(oracle, bank, v1, v2, idx, 3000, T, Z1, Z2, dtype=torch.bfloat16)
    for variant in (0, 2, 3):
        got = _score(pkg, cuda, bank, v1, v2, idx, 3000, T, Z1, Z2, dtype=torch.bfloat16, variant=variant)
        _check_full(got, want_same, REL32)
    # and vs the fp32-bank oracle: bf16 tolerance
    want32 = _oracle_score(oracle, bank, v1, v2, idx, 3000, T, Z1, Z2)
    assert _rel(got["res"][0] + got["res"][1], want32["loss_s"] + want32["loss_t"]) < RELBF
    assert _rel(got["g1"], want32["grad_v1"]) < RELBF


def _update(pkg, dev, bank, v1, v2, y, m, row_begin=0, row_end=None, dtype=torch.float32):
    lib = pkg._native.lib()
    N, _, D = bank.shape
    row_end = N if row_end is None else row_end
    dbank = bank[row_begin:row_end].to(dtype).contiguous().to(dev)
    dv1, dv2, dy = v1.to(dev), v2.to(dev), y.to(dev)
    m32 = float(np.float32(m))
    esz = 4 if dtype == torch.float32 else 2
    rc = lib.crdpn_crd_momentum_update(dbank.data_ptr(), dbank.data_ptr() + D * esz, 2 * D, 0 if esz == 4 else 1,
                                       dv1.data_ptr(), dv2.data_ptr(), dy.data_ptr(), len(y), D, row_begin, row_end,
                                       m32, 1.0 - m32, torch.cuda.current_stream().cuda_stream)
    pkg._native.check(rc, "update")
    torch.cuda.synchronize()
    return dbank.cpu()


@pytest.mark.parametrize("D", [32, 64, 128, 256, 512])
def test_momentum_update_bit_exact(pkg, oracle, cuda, D):
    bank, v1, v2, y, idx = _case(13, 4, D, 200, seed=D)
    y[5] = y[2]      # duplicate sample index: last occurrence wins
    y[11] = y[2]
    got = _update(pkg, cuda, bank, v1, v2, y, 0.5).numpy()
    want = bank.numpy().copy()
    oracle.momentum_update(want[:, 0, :], v1.numpy(), y.numpy(), 0.5)
    oracle.momentum_update(want[:, 1, :], v2.numpy(), y.numpy(), 0.5)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # vs the torch formula (different reduction order): within 2 ulp
    pos = bank[y, 0, :] * 0.5 + v1 * 0.5
    t = (pos / pos.pow(2).sum(1, keepdim=True).pow(0.5)).numpy()
    keep = [i for i in range(13) if i not in (2, 5)]
    assert np.max(np.abs(got[y.numpy()[keep], 0, :] - t[keep]) / np.spacing(np.abs(t[keep]))) <= 2


def test_momentum_update_shard_and_bf16(pkg, oracle, cuda):
    bank, v1, v2, y, idx = _case(16, 4, 128, 100, seed=2)
    got = _update(pkg, cuda, bank, v1, v2, y, 0.3, row_begin=40, row_end=90).numpy()
    want = bank[40:90].numpy().copy()
    oracle.momentum_update(want[:, 0, :], v1.numpy(), y.numpy(), 0.3, row_begin=40, row_end=90)
    oracle.momentum_update(want[:, 1, :], v2.numpy(), y.numpy(), 0.3, row_begin=40, row_end=90)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # bf16 banks: the fp32 canonical result rounded to nearest-even bf16
    gotb = _update(pkg, cuda, bank, v1, v2, y, 0.5, dtype=torch.bfloat16)
    wb = bank.to(torch.bfloat16).float().numpy().copy()
    oracle.momentum_update(wb[:, 0, :], v1.numpy(), y.numpy(), 0.5)
    oracle.momentum_update(wb[:, 1, :], v2.numpy(), y.numpy(), 0.5)
    assert torch.equal(gotb, torch.from_numpy(wb).to(torch.bfloat16))


def test_alias_draw_bit_exact(pkg, oracle, cuda):
    probs = torch.rand(1000, generator=torch.Generator().manual_seed(3)) ** 2
    am = pkg.AliasMethod(probs.clone(), seed=1234)
    prob, alias = oracle.alias_build(probs.numpy())
    assert np.array_equal(am.prob.numpy(), prob) and np.array_equal(am.alias.numpy(), alias)
    am.cuda()
    a = am.draw(100003).cpu().numpy()
    b = am.draw(77).cpu().numpy()
    assert np.array_equal(a, oracle.alias_draw(prob, alias, 100003, seed=1234, offset=0))
    assert np.array_equal(b, oracle.alias_draw(prob, alias, 77, seed=1234, offset=100003))
    y = torch.tensor([5, 999, 0, 17], device=cuda)
    c = am.draw_contrast(y, 1025).cpu().numpy()
    assert np.array_equal(c, oracle.alias_draw_contrast(prob, alias, y.cpu().numpy(), 1025, seed=1234, offset=100080))
    # a shard's sampler: local draws shifted by the shard's first row, column 0 still the GLOBAL positive index
    c2 = am.draw_contrast(y + 70000, 513, row_base=70000).cpu().numpy()
    want = oracle.alias_draw_contrast(prob, alias, y.cpu().numpy(), 513, seed=1234, offset=100080 + 4 * 1025)
    want[:, 1:] += 70000
    want[:, 0] = y.cpu().numpy() + 70000
    assert np.array_equal(c2, want)
    # uniform unigrams over a large N: every value in range, all residues reachable
    am2 = pkg.AliasMethod(torch.ones(90000), seed=7).cuda()
    d = am2.draw(1 << 20)
    assert int(d.min()) >= 0 and int(d.max()) < 90000 and d.unique().numel() > 89000


def _opt(**kw):
    base = dict(s_dim=200, t_dim=200, feat_dim=128, n_data=5000, nce_k=1024, nce_t=0.07, nce_m=0.5)
    base.update(kw)
    return type("Opt", (), base)()


def test_crdloss_module_matches_stock_formulation_two_steps(pkg, oracle, cuda):
    """Drop-in check: same weights and banks, fixed contrast_idx -> same loss, same gradients on inputs and
    embed parameters, same Z, bit-exact bank rows vs the canonical oracle, over two consecutive steps."""
    from oracle.crd_oracle import StockCRD
    opt = _opt()
    torch.manual_seed(46)
    crit = pkg.CRDLoss(opt).to(cuda)
    stock = StockCRD(opt.s_dim, opt.t_dim, opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m)
    with torch.no_grad():
        crit.embed_s.linear.weight.copy_(stock.Ws); crit.embed_s.linear.bias.copy_(stock.bs)
        crit.embed_t.linear.weight.copy_(stock.Wt); crit.embed_t.linear.bias.copy_(stock.bt)
        crit.contrast.memory_v1.copy_(stock.memory_v1); crit.contrast.memory_v2.copy_(stock.memory_v2)
    B = 46
    for step in range(2):
        g = torch.Generator().manual_seed(100 + step)
        f_s = torch.randn(B, opt.s_dim, generator=g)
        f_t = torch.randn(B, opt.t_dim, generator=g)
        y = torch.randperm(opt.n_data, generator=g)[:B]
        cidx = torch.randint(0, opt.n_data, (B, opt.nce_k + 1), generator=g)
        cidx[:, 0] = y
        canon1 = crit.contrast.memory_v1.detach().cpu().numpy().copy()
        canon2 = crit.contrast.memory_v2.detach().cpu().numpy().copy()
        fs_d = f_s.to(cuda).requires_grad_()
        ft_d = f_t.to(cuda).requires_grad_()
        crit.zero_grad()
        loss = crit(fs_d, ft_d, y.to(cuda), cidx.to(cuda))
        assert loss.dim() == 0 and loss.dtype == torch.float32
        (loss * 0.8).backward()
        fs_c, ft_c = f_s.clone().requires_grad_(), f_t.clone().requires_grad_()
        for p in (stock.Ws, stock.bs, stock.Wt, stock.bt):
            p.grad = None
        want = stock.loss(fs_c, ft_c, y, cidx)
        (want * 0.8).backward()
        assert _rel(loss.item(), want.item()) < REL32
        assert _rel(crit.contrast.params[2].item(), stock.Z1) < REL32 and _rel(crit.contrast.params[3].item(), stock.Z2) < REL32
        assert _rel(fs_d.grad.cpu(), fs_c.grad) < REL32 and _rel(ft_d.grad.cpu(), ft_c.grad) < REL32
        assert _rel(crit.embed_s.linear.weight.grad.cpu(), stock.Ws.grad) < REL32
        assert _rel(crit.embed_t.linear.bias.grad.cpu(), stock.bt.grad) < REL32
        # bank rows: bit-exact vs canonical oracle applied to the GPU's own embeddings; <=2 ulp vs torch
        with torch.no_grad():
            v1 = crit.embed_s(fs_d).cpu().numpy()
            v2 = crit.embed_t(ft_d).cpu().numpy()
        oracle.momentum_update(canon1, v1, y.numpy(), 0.5)
        oracle.momentum_update(canon2, v2, y.numpy(), 0.5)
        got1 = crit.contrast.memory_v1.detach().cpu().numpy()
        assert np.array_equal(got1.view(np.uint32), canon1.view(np.uint32))
        assert np.array_equal(crit.contrast.memory_v2.detach().cpu().numpy().view(np.uint32), canon2.view(np.uint32))
        assert _rel(got1, stock.memory_v1.numpy()) < 1e-5
        # keep the two implementations in lock-step for the next iteration
        with torch.no_grad():
            stock.memory_v1.copy_(torch.from_numpy(got1))
            stock.memory_v2.copy_(crit.contrast.memory_v2.detach().cpu())


def test_contrast_memory_forward_surface_and_internal_sampling(pkg, oracle, cuda):
    torch.manual_seed(1)
    mem = pkg.ContrastMemory(128, 3000, 511, seed=99).to(cuda)
    assert mem.memory_v1.stride(0) == 256 and mem.memory_v2.data_ptr() - mem.memory_v1.data_ptr() == 512
    v1 = torch.nn.functional.normalize(torch.randn(6, 128, device=cuda), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(6, 128, device=cuda), dim=1)
    y = torch.tensor([5, 17, 2999, 0, 44, 1000], device=cuda)
    b1 = mem.memory_v1.cpu().numpy().copy(); b2 = mem.memory_v2.cpu().numpy().copy()
    o1, o2 = mem(v1, v2, y)                      # idx=None -> on-device alias draw, column 0 <- y
    assert o1.shape == (6, 512, 1) and o2.shape == (6, 512, 1)
    prob, alias = oracle.alias_build(np.ones(3000, np.float32))
    cidx = oracle.alias_draw_contrast(prob, alias, y.cpu().numpy(), 512, seed=99, offset=0)
    Z1, Z2 = mem.params[2].item(), mem.params[3].item()
    want = oracle.crd_score(b1, b2, v1.cpu().numpy(), v2.cpu().numpy(), cidx, 3000, 0.07, Z1, Z2)
    assert _rel(o1.cpu().numpy()[:, :, 0], want["out_v1"]) < REL32
    assert _rel(o2.cpu().numpy()[:, :, 0], want["out_v2"]) < REL32
    # unfused criterion on those outputs == fused loss on the same inputs
    crit = pkg.ContrastLoss(3000)
    unfused = (crit(o1) + crit(o2)).item()
    assert _rel(unfused, want["loss_s"] + want["loss_t"]) < REL32


def test_config1_full_size_vs_oracle(pkg, oracle, cuda):
    """BASELINE config 1 at full size: B=46, D=128, K=16384, N=90k x 2, tau=0.07 (oracle takes a few seconds)."""
    B, K1, D, N = 46, 16385, 128, 90000
    bank, v1, v2, y, idx = _case(B, K1, D, N, seed=46)
    r0 = _score(pkg, cuda, bank, v1, v2, idx, N, 0.07, -1.0, -1.0, want_out=False)
    Z1 = float(np.float32(r0["res"][2] / r0["res"][4] * N))
    Z2 = float(np.float32(r0["res"][3] / r0["res"][4] * N))
    got = _score(pkg, cuda, bank, v1, v2, idx, N, 0.07, Z1, Z2, want_out=False)
    want = _oracle_score(oracle, bank, v1, v2, idx, N, 0.07, Z1, Z2)
    _check_full(got, want, REL32)


def test_fused_step_equals_score_then_update_bitwise(pkg, cuda):
    """crdpn_crd_step (2 launches) == crdpn_crd_score + crdpn_crd_momentum_update (3 launches), bit for bit."""
    import copy
    torch.manual_seed(5)
    a = pkg.ContrastMemory(128, 7000, 2048, seed=1).to(cuda)
    b = copy.deepcopy(a)
    assert b.memory_v1.stride(0) == 256   # deep copy keeps the interleaved allocation
    v1 = torch.nn.functional.normalize(torch.randn(46, 128, device=cuda), dim=1)
    v2 = torch.nn.functional.normalize(torch.randn(46, 128, device=cuda), dim=1)
    y = torch.randperm(7000, device=cuda)[:46]
    y[7] = y[3]
    idx = torch.randint(0, 7000, (46, 2049), device=cuda)
    idx[:, 0] = y
    res_a, g1a, g2a = a._step(v1, v2, y, idx, 1234.5, 2345.5)
    res_a = res_a.clone()
    res_b, g1b, g2b, _, _ = b._score(v1, v2, idx, 1234.5, 2345.5)
    b._update(v1, v2, y)
    assert torch.equal(res_a, res_b) and torch.equal(g1a, g1b) and torch.equal(g2a, g2b)
    assert torch.equal(a.memory_v1, b.memory_v1) and torch.equal(a.memory_v2, b.memory_v2)
    assert not torch.equal(a.memory_v1[y[0]], torch.zeros(128, device=cuda))


@pytest.mark.parametrize("B,dim_in,D", [(46, 2048, 128), (46, 200, 128), (138, 1024, 128), (5, 37, 32), (64, 513, 256)])
def test_fused_embed_head_matches_eager_linear_l2norm(pkg, cuda, B, dim_in, D):
    """Embed = Linear + x/||x||_2 (published crd/criterion.py Embed/Normalize): fused kernels vs the eager
    sub-modules of the same module, forward and all three gradients, <= 1e-4 relative (fp32)."""
    torch.manual_seed(B + dim_in)
    emb = pkg.Embed(dim_in, D).to(cuda)
    x = torch.randn(B, dim_in, device=cuda)
    g = torch.randn(B, D, device=cuda)
    xa = x.clone().requires_grad_()
    va = emb(xa)                                                  # fused path
    va.backward(g)
    got = (va.detach(), xa.grad.clone(), emb.linear.weight.grad.clone(), emb.linear.bias.grad.clone())
    emb.zero_grad()
    xb = x.clone().requires_grad_()
    vb = emb.l2norm(emb.linear(xb))                               # eager sub-modules
    vb.backward(g)
    want = (vb.detach(), xb.grad, emb.linear.weight.grad, emb.linear.bias.grad)
    for a, b in zip(got, want):
        assert ((a - b).abs().max() / b.abs().max()).item() < 1e-4
    assert torch.allclose(va.detach().norm(dim=1), torch.ones(B, device=cuda), atol=1e-5)
    # no input gradient requested -> the dgrad kernel is skipped and autograd still gets the parameter gradients
    emb.zero_grad()
    emb(x).backward(g)
    assert ((emb.linear.weight.grad - want[2]).abs().max() / want[2].abs().max()).item() < 1e-4


@pytest.mark.parametrize("dup", [False, True])
def test_unfused_published_composition_is_differentiable(pkg, oracle, cuda, dup):
    """ContrastLoss(ContrastMemory(v1, v2, y, idx)) -- the published CRDLoss body, written out -- gives the stock
    formulation's loss AND gradients over two consecutive steps (the second step's contrast list deliberately hits rows
    the first step's momentum update rewrote, and rows of its own positives, which the backward must see PRE-update)."""
    from oracle.crd_oracle import StockCRD
    opt = _opt()
    torch.manual_seed(46)
    crit = pkg.CRDLoss(opt).to(cuda)
    stock = StockCRD(opt.s_dim, opt.t_dim, opt.feat_dim, opt.n_data, opt.nce_k, opt.nce_t, opt.nce_m)
    with torch.no_grad():
        crit.embed_s.linear.weight.copy_(stock.Ws); crit.embed_s.linear.bias.copy_(stock.bs)
        crit.embed_t.linear.weight.copy_(stock.Wt); crit.embed_t.linear.bias.copy_(stock.bt)
        crit.contrast.memory_v1.copy_(stock.memory_v1); crit.contrast.memory_v2.copy_(stock.memory_v2)
    B = 24
    y_prev = None
    for step in range(2):
        g = torch.Generator().manual_seed(300 + step)
        f_s = torch.randn(B, opt.s_dim, generator=g)
        f_t = torch.randn(B, opt.t_dim, generator=g)
        y = torch.randperm(opt.n_data, generator=g)[:B]
        if dup:
            y[B // 2:] = y[:B // 2]                      # duplicate positives inside the batch
        cidx = torch.randint(0, opt.n_data, (B, opt.nce_k + 1), generator=g)
        cidx[:, 0] = y
        cidx[:, 5] = y.roll(1)                           # negatives that are other anchors' positives (rewritten this step)
        if y_prev is not None:
            cidx[:, 7] = y_prev                          # rows rewritten by the previous step
        y_prev = y
        fs_d, ft_d = f_s.to(cuda).requires_grad_(), f_t.to(cuda).requires_grad_()
        crit.zero_grad()
        out_s, out_t = crit.contrast(crit.embed_s(fs_d), crit.embed_t(ft_d), y.to(cuda), cidx.to(cuda))
        assert out_s.shape == (B, opt.nce_k + 1, 1) and out_s.requires_grad
        loss = (crit.criterion_s(out_s) + crit.criterion_t(out_t)).squeeze()
        (loss * 0.6).backward()
        fs_c, ft_c = f_s.clone().requires_grad_(), f_t.clone().requires_grad_()
        for p in (stock.Ws, stock.bs, stock.Wt, stock.bt):
            p.grad = None
        want = stock.loss(fs_c, ft_c, y, cidx)
        (want * 0.6).backward()
        assert _rel(loss.item(), want.item()) < REL32
        assert _rel(fs_d.grad.cpu(), fs_c.grad) < REL32 and _rel(ft_d.grad.cpu(), ft_c.grad) < REL32
        assert _rel(crit.embed_s.linear.weight.grad.cpu(), stock.Ws.grad) < REL32
        assert _rel(crit.embed_t.linear.weight.grad.cpu(), stock.Wt.grad) < REL32
        if not dup:  # (with duplicate positives torch's index_copy_ order is the oracle's "last wins" -- covered elsewhere)
            assert _rel(crit.contrast.memory_v1.cpu(), stock.memory_v1) < 1e-6


def test_alias_uniform_shortcut_draws_the_same_indices(pkg, oracle, cuda):
    """Uniform unigrams build prob == 1: the draw skips the table gather and must yield the oracle's indices anyway;
    a skewed distribution keeps using the tables."""
    N = 5000
    am = pkg.AliasMethod(torch.ones(N), seed=77).cuda()
    assert am.uniform and am.table_ptrs() == (None, None)
    got = am.draw(4096).cpu().numpy()
    prob, alias = oracle.alias_build(np.ones(N, dtype=np.float32))
    assert np.array_equal(got, oracle.alias_draw(prob, alias, 4096, 77, 0))
    skew = pkg.AliasMethod(torch.arange(1, N + 1, dtype=torch.float32), seed=77).cuda()
    assert not skew.uniform
    prob, alias = oracle.alias_build(np.arange(1, N + 1, dtype=np.float32))
    assert np.array_equal(skew.draw(4096).cpu().numpy(), oracle.alias_draw(prob, alias, 4096, 77, 0))


def test_headline_config_properties_full_size(pkg, oracle, cuda):
    """BASELINE configs[3] on one GPU (B=46, D=128, K=65536, N=1M rows x 2 banks) -- far too big for the scalar oracle as a
    whole, so parity is carried by properties that do not depend on the size:
      * two anchors scored by the oracle (their K+1 entries against the full 1M-row banks): gradient rows and the
        anchors' loss terms equal the kernel's (a gradient row depends on its own anchor only);
      * anchor additivity: loss(all 46) = sum_b loss(anchor b alone) / 46;
      * shard additivity: four row shards' losses, counts and gradients add up to the unsharded result;
      * permuting the K negative columns changes nothing beyond fp32 summation order;
      * bit-reproducibility run to run."""
    B, K1, D, N, T, Z1, Z2 = 46, 65537, 128, 1_000_000, 0.07, 2.1e6, 2.2e6
    bank_cpu, v1, v2, y, idx = _case(B, K1, D, N, seed=46)
    bank = bank_cpu.to(cuda)          # 1 GB, uploaded once
    full = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, want_out=False)
    again = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, want_out=False)
    assert np.array_equal(full["res"], again["res"]) and np.array_equal(full["g1"], again["g1"])
    assert full["res"][4] == B * K1
    # (1) oracle on two anchors
    sh = bank_cpu.numpy()
    tot_s = tot_t = 0.0
    single = {}
    for b in range(B):
        one = _score(pkg, cuda, bank, v1[b:b + 1], v2[b:b + 1], idx[b:b + 1], N, T, Z1, Z2, want_out=False)
        single[b] = one
        tot_s += one["res"][0] / B
        tot_t += one["res"][1] / B
        # a one-anchor batch scales dL/ds by 1/(1*T) instead of 1/(46*T)
        assert _rel(one["g1"][0] / B, full["g1"][b]) < 1e-5 and _rel(one["g2"][0] / B, full["g2"][b]) < 1e-5
    for b in (0, 45):
        want = oracle.crd_score(sh[:, 0, :], sh[:, 1, :], v1[b:b + 1].numpy(), v2[b:b + 1].numpy(), idx[b:b + 1].numpy(),
                                N, T, Z1, Z2, want_out=False)
        assert _rel(single[b]["res"][0], want["loss_s"]) < REL32 and _rel(single[b]["res"][1], want["loss_t"]) < REL32
        assert _rel(single[b]["g1"], want["grad_v1"]) < REL32 and _rel(single[b]["g2"], want["grad_v2"]) < REL32
    # (2) anchor additivity
    assert _rel(tot_s, full["res"][0]) < 1e-6 and _rel(tot_t, full["res"][1]) < 1e-6
    # (3) shard additivity
    acc = dict(ls=0.0, lt=0.0, cnt=0.0, g1=np.zeros((B, D)), g2=np.zeros((B, D)))
    for r in range(4):
        lo, hi = N * r // 4, N * (r + 1) // 4
        part = _score(pkg, cuda, bank, v1, v2, idx, N, T, Z1, Z2, row_begin=lo, row_end=hi, want_out=False)
        acc["ls"] += part["res"][0]; acc["lt"] += part["res"][1]; acc["cnt"] += part["res"][4]
        acc["g1"] += part["g1"]; acc["g2"] += part["g2"]
    assert acc["cnt"] == B * K1
    assert _rel(acc["ls"], full["res"][0]) < 1e-6 and _rel(acc["lt"], full["res"][1]) < 1e-6
    assert _rel(acc["g1"], full["g1"]) < 1e-5 and _rel(acc["g2"], full["g2"]) < 1e-5
    # (4) column permutation
    perm = torch.randperm(K1 - 1, generator=torch.Generator().manual_seed(1)) + 1
    idx_p = idx.clone()
    idx_p[:, 1:] = idx[:, perm]
    p = _score(pkg, cuda, bank, v1, v2, idx_p, N, T, Z1, Z2, want_out=False)
    assert _rel(p["res"][0], full["res"][0]) < 1e-6 and _rel(p["g1"], full["g1"]) < 1e-5 and _rel(p["g2"], full["g2"]) < 1e-5


def test_int32_contrast_idx_and_in_kernel_draw_are_bit_identical_to_the_int64_list(pkg, cuda):
    """Three sources of the same contrast indices -- an int64 list, the same list as int32, and no list at all (the scoring
    pass draws the entries itself from the sampler's Philox stream) -- must give the same bits: loss, gradients, bank rows."""
    opt = type("Opt", (), dict(s_dim=64, t_dim=48, feat_dim=128, n_data=5000, nce_k=1500, nce_t=0.07, nce_m=0.5))()
    g = torch.Generator().manual_seed(9)
    B = 23
    f_s, f_t = torch.randn(B, 64, generator=g).to(cuda), torch.randn(B, 48, generator=g).to(cuda)
    y = torch.randperm(5000, generator=g)[:B].to(cuda)
    mods = []
    for _ in range(3):
        torch.manual_seed(4)
        mods.append(pkg.CRDLoss(opt, seed=1234).to(cuda))
    outs = []
    for mode, m in zip(("drawn", "int64", "int32"), mods):
        rows = []
        for step in range(3):   # step 0 freezes Z (general path), steps 1-2 run the one-call path
            if mode == "drawn":
                cidx = None
            else:
                smp = pkg.AliasMethod(torch.ones(5000), seed=1234).cuda()
                smp.offset = step * B * 1501
                cidx = smp.draw_contrast(y, 1501)
                if mode == "int32":
                    cidx = cidx.to(torch.int32)
            fs = f_s.clone().requires_grad_()
            loss = m(fs, f_t, y, cidx)
            loss.backward()
            rows.append((loss.detach().clone(), fs.grad.clone(), m.embed_t.linear.weight.grad.clone(), m.contrast.memory_v1.clone()))
            m.zero_grad()
        outs.append(rows)
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            for ta, tb in zip(a, b):
                assert torch.equal(ta, tb)


def test_step_pipeline_reads_every_loss_and_changes_no_bits(pkg, cuda):
    """StepPipeline (staged H2D copies on a copy stream, losses read one step late) against the reference loop verbatim
    (copy, forward, backward, loss.item()): the same losses, bit for bit, and the same bank contents afterwards."""
    opt = type("Opt", (), dict(s_dim=64, t_dim=48, feat_dim=128, n_data=4000, nce_k=1000, nce_t=0.07, nce_m=0.5))()
    g = torch.Generator().manual_seed(3)
    B, steps = 16, 6
    batches = [(torch.randn(B, 64, generator=g).pin_memory(), torch.randn(B, 48, generator=g).pin_memory(),
                torch.randperm(4000, generator=g)[:B].pin_memory()) for _ in range(steps)]
    mods = []
    for _ in range(2):
        torch.manual_seed(8)
        mods.append(pkg.CRDLoss(opt, seed=99).to(cuda))

    def fwd_bwd(m, dev_in):
        f_s, f_t, y = dev_in
        f_s.requires_grad_()
        loss = m(f_s, f_t, y)
        m.zero_grad(set_to_none=True)
        loss.backward()
        return loss

    strict = [fwd_bwd(mods[0], [t.to(cuda, non_blocking=True) for t in b]).item() for b in batches]
    pipe, got = pkg.StepPipeline(cuda), []
    pipe.stage(*batches[0])
    for i in range(steps):
        dev_in = pipe.take()
        if i + 1 < steps:
            pipe.stage(*batches[i + 1])
        pipe.publish(fwd_bwd(mods[1], dev_in))
        if pipe.pending() > 1:
            got.append(pipe.collect())
    while pipe.pending():
        got.append(pipe.collect())
    assert got == strict
    assert torch.equal(mods[0].contrast.memory_v1, mods[1].contrast.memory_v1)
    assert torch.equal(mods[0].contrast.memory_v2, mods[1].contrast.memory_v2)


def test_graphed_step_equals_the_eager_loop(pkg, cuda):
    """GraphedStep: forward + backward of CRDLoss(f_s, f_t, idx) captured once and replayed from staged host batches must
    reproduce the eager loop step by step -- fresh negatives on every replay (device-resident sampler offset), the same
    losses, gradients and bank rows, bit for bit."""
    B, D, K, N = 12, 128, 511, 6000
    opt = type("Opt", (), dict(s_dim=96, t_dim=64, feat_dim=D, n_data=N, nce_k=K, nce_t=0.07, nce_m=0.5))()
    torch.manual_seed(7)
    a = pkg.CRDLoss(opt, seed=11).to(cuda)
    torch.manual_seed(7)
    b = pkg.CRDLoss(opt, seed=11).to(cuda)
    b.load_state_dict(a.state_dict())
    gen = torch.Generator().manual_seed(3)
    batches = [(torch.randn(B, 96, generator=gen).pin_memory(), torch.randn(B, 64, generator=gen).pin_memory(),
                torch.randperm(N, generator=gen)[:B].pin_memory()) for _ in range(5)]
    # first call freezes Z on both, eagerly, from the same batch
    for m in (a, b):
        f_s = batches[0][0].to(cuda).requires_grad_(True)
        m(f_s, batches[0][1].to(cuda), batches[0][2].to(cuda)).backward()
        m.zero_grad(set_to_none=True)
    b.contrast.device_sampler_offset()

    def fwd_bwd(f_s, f_t, idx):
        loss = b(f_s, f_t, idx)
        loss.backward()
        return loss

    # the warm-up replays inside GraphedStep advance the sampler and the banks: give the eager module the same history
    warm = 2                      # (the capture pass itself executes nothing)
    step = pkg.GraphedStep(fwd_bwd, batches[0], cuda, grad_inputs=(0,), zero_grad=lambda: b.zero_grad(set_to_none=True), warmup=warm)
    for _ in range(warm):
        f_s = batches[0][0].to(cuda).requires_grad_(True)
        a(f_s, batches[0][1].to(cuda), batches[0][2].to(cuda)).backward()
        a.zero_grad(set_to_none=True)
    step.stage(*batches[1])
    for i in range(1, 5):
        step.run()
        if i + 1 < 5:
            step.stage(*batches[i + 1])
        f_s = batches[i][0].to(cuda).requires_grad_(True)
        a.zero_grad(set_to_none=True)
        want = a(f_s, batches[i][1].to(cuda), batches[i][2].to(cuda))
        want.backward()
        got = step.collect()
        assert got == want.item(), (i, got, want.item())
        assert torch.equal(step.static[0].grad, f_s.grad)
        assert torch.equal(b.embed_s.linear.weight.grad, a.embed_s.linear.weight.grad)
        assert torch.equal(b.embed_t.linear.bias.grad, a.embed_t.linear.bias.grad)
    assert torch.equal(b.contrast.memory_v1, a.contrast.memory_v1) and torch.equal(b.contrast.memory_v2, a.contrast.memory_v2)
    b.contrast.device_sampler_offset(False)
    assert b.contrast.multinomial.offset == a.contrast.multinomial.offset


def test_band_sorted_step_matches_the_gather_step(pkg, oracle, cuda):
    """ContrastMemory.sweep (variant | 0x400): every (anchor, chunk) list stably sorted by row band, the warps of a unit taking
    interleaved blocks of it.  Same scores, so: loss / gradients equal the gather step's up to fp32 summation order, updated
    rows bit-identical, bit-reproducible run to run, and (two anchors) equal to the oracle.  Also through CRDLoss with the
    negatives drawn inside the pre-pass, and with an int32 list."""
    B, D, K, N, T = 46, 128, 65536, 20000, 0.07
    gen = torch.Generator().manual_seed(12)
    bank = torch.nn.functional.normalize(torch.randn(N, 2, D, generator=gen), dim=2)
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen)).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen)).to(cuda)
    y = torch.randperm(N, generator=gen)[:B].to(cuda)
    cidx = torch.randint(0, N, (B, K + 1), generator=gen)
    cidx[:, 0] = y.cpu()
    Z1, Z2 = 4.1e4, 4.3e4
    runs = {}
    for name, sweep, lst in (("gather", False, cidx), ("swept", True, cidx), ("swept_again", True, cidx), ("swept_i32", True, cidx.to(torch.int32))):
        mem = pkg.ContrastMemory(D, N, K, T, 0.5).to(cuda)
        with torch.no_grad():
            mem.memory_v1.copy_(bank[:, 0]); mem.memory_v2.copy_(bank[:, 1]); mem.params[2], mem.params[3] = Z1, Z2
        mem._host = None
        mem.sweep = sweep
        if lst.dtype == torch.int32:
            mem.variant = mem.IDX32       # (ContrastMemory._step takes the list as it is; CRDLoss sets this bit itself)
        assert bool(mem._step_variant(B, K + 1, D) & mem.SWEEP) == sweep
        res, g1, g2 = mem._step(v1, v2, y, lst.to(cuda), Z1, Z2)
        torch.cuda.synchronize()
        runs[name] = (res.clone(), g1.clone(), g2.clone(), mem.memory_v1[y].clone(), mem.memory_v2[y].clone())
    a, b = runs["gather"], runs["swept"]
    assert abs((a[0][5] - b[0][5]).item()) <= 1e-6 * abs(a[0][5].item()) and a[0][4].item() == b[0][4].item() == B * (K + 1)
    assert _rel(b[1].cpu().numpy(), a[1].cpu().numpy()) < 1e-5 and _rel(b[2].cpu().numpy(), a[2].cpu().numpy()) < 1e-5
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    for other in ("swept_again", "swept_i32"):
        assert all(torch.equal(p, q) for p, q in zip(b, runs[other])), other
    sub = [0, B - 1]
    want = oracle.crd_score(bank[:, 0].contiguous().numpy(), bank[:, 1].contiguous().numpy(), v1[sub].cpu().numpy(), v2[sub].cpu().numpy(),
                            cidx[sub].numpy(), N, T, Z1, Z2)
    scale = B / len(sub)
    assert _rel(b[1][sub].cpu().numpy(), want["grad_v1"] / scale) < REL32 and _rel(b[2][sub].cpu().numpy(), want["grad_v2"] / scale) < REL32
    # negatives drawn on the GPU (no list in memory): the pre-pass draws them itself, the same ones as the gather step
    opt = type("Opt", (), dict(s_dim=64, t_dim=48, feat_dim=D, n_data=N, nce_k=K, nce_t=T, nce_m=0.5))()
    losses = []
    for sweep in (False, True):
        torch.manual_seed(4)
        crit = pkg.CRDLoss(opt, seed=9).to(cuda)
        crit.contrast.sweep = sweep
        f_s = torch.randn(B, 64, generator=torch.Generator().manual_seed(1)).to(cuda).requires_grad_(True)
        f_t = torch.randn(B, 48, generator=torch.Generator().manual_seed(2)).to(cuda)
        for _ in range(2):
            loss = crit(f_s, f_t, y)
        loss.backward()
        losses.append((loss.item(), f_s.grad.clone(), crit.contrast.memory_v1[y].clone()))
    assert abs(losses[0][0] - losses[1][0]) <= 1e-6 * abs(losses[0][0])
    assert _rel(losses[1][1].cpu().numpy(), losses[0][1].cpu().numpy()) < 1e-5 and torch.equal(losses[0][2], losses[1][2])


@pytest.mark.parametrize("B,K,N,lo,hi", [(7, 300, 5000, 0, 5000), (46, 2048, 6000, 0, 6000), (512, 4096, 9000, 0, 9000),
                                         (138, 16384, 20000, 0, 20000), (46, 70001, 30000, 0, 30000),
                                         (46, 65536, 40000, 10000, 20000), (30, 131072, 8192, 0, 8192), (3, 1, 4096, 0, 4096)])
def test_band_sorted_partition_over_shapes(pkg, cuda, B, K, N, lo, hi):
    """The band-sorted mode over the shapes that stress its partition: one chunk per anchor, hundreds of anchors (few warps
    each), 32 chunks per anchor, K + 1 not a multiple of the chunk, a shard in the middle of the bank (survivor compaction),
    tripled indices, a single negative.  Against the plain step on the same inputs: the entry count is exact (nothing lost or
    scored twice), loss / gradients agree to fp32 summation order, updated rows are bit-identical."""
    D, T, Z1, Z2 = 128, 0.07, 3.0e4, 3.1e4
    gen = torch.Generator().manual_seed(B * 131 + K)
    rows = hi - lo
    bank = torch.nn.functional.normalize(torch.randn(rows, 2, D, generator=gen), dim=2)
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen)).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=gen)).to(cuda)
    y = torch.randint(0, N, (B,), generator=gen)
    if B == 138:
        y = y[:46].repeat(3)
    cidx = torch.randint(0, N, (B, K + 1), generator=gen)
    cidx[:, 0] = y
    y, cidx = y.to(cuda), cidx.to(cuda)
    out = {}
    for name, sweep in (("plain", False), ("swept", True)):
        mem = pkg.ContrastMemory(D, N, K, T, 0.5, row_begin=lo, row_end=hi).to(cuda)
        with torch.no_grad():
            mem.memory_v1.copy_(bank[:, 0]); mem.memory_v2.copy_(bank[:, 1]); mem.params[2], mem.params[3] = Z1, Z2
        mem._host = None
        mem.sweep = sweep
        l0 = pkg._native.launch_count()
        res, g1, g2 = mem._step(v1, v2, y, cidx, Z1, Z2)
        torch.cuda.synchronize()
        assert pkg._native.launch_count() - l0 == (3 if sweep else 2)   # the band-sort pre-pass really ran (no silent fall-back)
        out[name] = (res.cpu().numpy().copy(), g1.cpu().numpy().copy(), g2.cpu().numpy().copy(),
                     mem.memory_v1.clone(), mem.memory_v2.clone())
    a, b = out["plain"], out["swept"]
    want_cnt = int(((cidx >= lo) & (cidx < hi)).sum().item())
    assert a[0][4] == b[0][4] == want_cnt
    assert abs(a[0][5] - b[0][5]) <= 2e-6 * abs(a[0][5]) + 1e-12
    assert _rel(b[1], a[1]) < 2e-5 and _rel(b[2], a[2]) < 2e-5
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
