"""CPU tests that freeze the CRD oracle (oracle/crd_oracle.{c,py}).

The reference ships no CRD code, tests or vectors (parity unpinned), so the pins are
known-answer cases worked out here independently of the oracle: a hand-computable N=4,K=2,D=2 case in
plain Python floats, Vose tables for a skewed 5-element distribution, the Philox4x32-10 known-answer
vector, Z-freeze across two calls, update-then-score ordering, and agreement between the two independent
restatements (closed-form C vs. autograd over the stock torch formulation).
"""
import math

import numpy as np
import pytest
import torch


def test_philox_known_answer(oracle):
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    out = oracle.philox(0, 0)
    assert [hex(int(v)) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


def _vose_python(probs):
    """Independent restatement in numpy-fp32 scalars of the published stack pairing."""
    p = np.asarray(probs, dtype=np.float32)
    n = len(p)
    if p.sum(dtype=np.float64).astype(np.float32) > 1:
        p = p / p.sum(dtype=np.float64).astype(np.float32)
    prob = np.zeros(n, np.float32)
    alias = np.zeros(n, np.int64)
    small, large = [], []
    for k in range(n):
        prob[k] = np.float32(n) * p[k]
        (small if prob[k] < 1.0 else large).append(k)
    while small and large:
        s, l = small.pop(), large.pop()
        alias[s] = l
        prob[l] = (prob[l] - np.float32(1.0)) + prob[s]
        (small if prob[l] < 1.0 else large).append(l)
    for k in small + large:
        prob[k] = 1
    return prob, alias


def test_alias_tables_skewed_known_answer(oracle):
    probs = np.array([0.1, 0.2, 0.3, 0.15, 0.25], dtype=np.float32)
    prob, alias = oracle.alias_build(probs)
    # hand trace: smaller=[0,3], larger=[1,2,4] -> (3,4) -> (0,4) -> (4,2)
    assert alias.tolist() == [4, 0, 0, 4, 2]
    np.testing.assert_allclose(prob, [0.5, 1.0, 1.0, 0.75, 0.5], atol=1e-6)
    implied = prob.astype(np.float64).copy()
    for j in range(5):
        implied[alias[j]] += 1.0 - prob[j]
    np.testing.assert_allclose(implied / 5, probs, atol=1e-6)
    p2, a2 = _vose_python(probs)
    assert np.array_equal(prob, p2) and np.array_equal(alias, a2)


@pytest.mark.parametrize("n", [7, 1000, 90000])
def test_alias_tables_uniform_degenerate(oracle, n):
    prob, alias = oracle.alias_build(np.ones(n, dtype=np.float32))
    # sum > 1 -> normalised; N*(1/N) == 1 in fp32 for these N -> every prob 1, alias 0
    p2, a2 = _vose_python(np.ones(n, dtype=np.float32))
    assert np.array_equal(prob, p2) and np.array_equal(alias, a2)
    if n in (1000, 90000):
        assert np.all(prob == 1.0) and np.all(alias == 0)


def test_alias_tables_random_vs_python(oracle):
    rng = np.random.default_rng(3)
    probs = rng.random(257).astype(np.float32) ** 3
    prob, alias = oracle.alias_build(probs)
    p2, a2 = _vose_python(probs)
    assert np.array_equal(prob, p2) and np.array_equal(alias, a2)


def test_alias_draw_properties(oracle):
    probs = np.array([0.1, 0.2, 0.3, 0.15, 0.25], dtype=np.float32)
    prob, alias = oracle.alias_build(probs)
    a = oracle.alias_draw(prob, alias, 200000, seed=46)
    assert a.min() >= 0 and a.max() < 5
    freq = np.bincount(a, minlength=5) / a.size
    np.testing.assert_allclose(freq, probs, atol=5e-3)
    # counter-based: a draw continued at offset equals the tail of a longer draw; seed changes the stream
    b = oracle.alias_draw(prob, alias, 1000, seed=46, offset=1000)
    assert np.array_equal(b, a[1000:2000])
    assert not np.array_equal(oracle.alias_draw(prob, alias, 1000, seed=47), a[:1000])
    y = np.array([3, 1, 4], dtype=np.int64)
    c = oracle.alias_draw_contrast(prob, alias, y, 6, seed=46)
    assert c.shape == (3, 6) and np.array_equal(c[:, 0], y)
    assert np.array_equal(c.reshape(-1)[1:6], a[1:6])


def test_tiny_hand_computed_case(oracle):
    # N=4, K=2, D=2, B=1: everything by hand in Python floats
    bank1 = np.array([[1.0, 0.0], [0.0, 1.0], [0.6, 0.8], [-1.0, 0.0]], np.float32)
    bank2 = np.array([[0.0, 1.0], [1.0, 0.0], [0.8, 0.6], [0.0, -1.0]], np.float32)
    v1 = np.array([[0.6, 0.8]], np.float32)
    v2 = np.array([[0.8, -0.6]], np.float32)
    idx = np.array([[2, 0, 3]], np.int64)
    T, N, Z1, Z2, eps = 0.5, 4, 3.0, 5.0, 1e-7
    f = lambda a: float(np.float32(a))
    s1 = [f(0.8) * f(0.6) + f(0.6) * f(0.8), f(0.8), -f(0.8)]           # bank2[idx] . v1
    s2 = [f(0.6) * f(0.8) - f(0.8) * f(0.6), f(0.8), -f(0.8)]           # bank1[idx] . v2
    o1 = [math.exp(s / T) / Z1 for s in s1]
    o2 = [math.exp(s / T) / Z2 for s in s2]
    # the published ContrastLoss adds / fills these two Python floats into fp32 tensors: they act as float32 values
    mPn = f(2 / N)
    c = f(2 / N + eps)
    nce = lambda o: -(math.log(o[0] / (o[0] + c)) + sum(math.log(mPn / (x + c)) for x in o[1:])) / 1
    res = oracle.crd_score(bank1, bank2, v1, v2, idx, N, T, Z1, Z2, eps)
    assert res["loss_s"] == pytest.approx(nce(o1), rel=1e-12)
    assert res["loss_t"] == pytest.approx(nce(o2), rel=1e-12)
    np.testing.assert_allclose(res["out_v1"][0], o1, rtol=1e-12)
    np.testing.assert_allclose(res["out_v2"][0], o2, rtol=1e-12)
    d1 = [-c / (T * (o1[0] + c))] + [x / (T * (x + c)) for x in o1[1:]]
    g1 = sum(d * bank2[i].astype(np.float64) for d, i in zip(d1, idx[0]))
    np.testing.assert_allclose(res["grad_v1"][0], g1, rtol=1e-12)
    # sum mode (Z unset): raw exponentials and their sums, no loss
    res0 = oracle.crd_score(bank1, bank2, v1, v2, idx, N, T, -1.0, -1.0, eps)
    assert res0["sum_e1"] == pytest.approx(sum(math.exp(s / T) for s in s1), rel=1e-12)
    assert res0["count"] == 3 and res0["loss_s"] == 0


def _setup(B=6, K=64, D=32, N=500, s_dim=40, t_dim=24, seed=5):
    from oracle.crd_oracle import StockCRD
    torch.manual_seed(seed)
    stock = StockCRD(s_dim, t_dim, D, N, K, 0.07, 0.5, seed=seed)
    f_s, f_t = torch.randn(B, s_dim), torch.randn(B, t_dim)
    y = torch.randperm(N)[:B]
    cidx = torch.randint(0, N, (B, K + 1))
    cidx[:, 0] = y
    return stock, f_s, f_t, y, cidx


def test_closed_form_matches_autograd_of_stock_formulation(oracle):
    stock, f_s, f_t, y, cidx = _setup()
    v1 = stock.embed(f_s, stock.Ws, stock.bs).detach().requires_grad_()
    v2 = stock.embed(f_t, stock.Wt, stock.bt).detach().requires_grad_()
    b1, b2 = stock.memory_v1.numpy().copy(), stock.memory_v2.numpy().copy()
    o1, o2 = stock.contrast(v1, v2, y, cidx)           # freezes Z, updates banks
    loss = (stock.nce(o1) + stock.nce(o2)).reshape(())
    loss.backward()
    res = oracle.crd_score(b1, b2, v1.detach().numpy(), v2.detach().numpy(), cidx.numpy(), stock.n_data,
                           stock.T, stock.Z1, stock.Z2)
    assert res["loss_s"] + res["loss_t"] == pytest.approx(loss.item(), rel=2e-5)
    np.testing.assert_allclose(res["grad_v1"], v1.grad.numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(res["grad_v2"], v2.grad.numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(res["out_v1"], o1.detach().numpy()[:, :, 0], rtol=2e-5)
    # first-call Z = mean(e) * N
    r0 = oracle.crd_score(b1, b2, v1.detach().numpy(), v2.detach().numpy(), cidx.numpy(), stock.n_data, stock.T, -1, -1)
    assert r0["sum_e1"] / r0["count"] * stock.n_data == pytest.approx(stock.Z1, rel=1e-5)
    assert r0["sum_e2"] / r0["count"] * stock.n_data == pytest.approx(stock.Z2, rel=1e-5)


def test_z_is_frozen_after_first_call():
    stock, f_s, f_t, y, cidx = _setup()
    stock.step(f_s, f_t, y, cidx)
    z = (stock.Z1, stock.Z2)
    assert z[0] > 0 and z[1] > 0
    stock.step(torch.randn_like(f_s), torch.randn_like(f_t), y, cidx)
    assert (stock.Z1, stock.Z2) == z


def test_scores_read_pre_update_bank_then_update(oracle):
    """A positive row that is also another anchor's negative must be scored with its pre-update value."""
    stock, f_s, f_t, y, cidx = _setup()
    cidx[1, 5] = y[0]
    b1 = stock.memory_v1.numpy().copy()
    b2 = stock.memory_v2.numpy().copy()
    with torch.no_grad():
        v1 = stock.embed(f_s, stock.Ws, stock.bs)
        v2 = stock.embed(f_t, stock.Wt, stock.bt)
        o1, _ = stock.contrast(v1, v2, y, cidx)
    res = oracle.crd_score(b1, b2, v1.numpy(), v2.numpy(), cidx.numpy(), stock.n_data, stock.T, stock.Z1, stock.Z2)
    np.testing.assert_allclose(res["out_v1"], o1.numpy()[:, :, 0], rtol=2e-5)
    assert not np.array_equal(stock.memory_v1.numpy()[y[0]], b1[y[0]])  # and the row did move afterwards


def test_momentum_update_canonical_vs_torch(oracle):
    stock, f_s, f_t, y, cidx = _setup(B=9, D=128)
    with torch.no_grad():
        v1 = stock.embed(f_s, stock.Ws, stock.bs)
    bank = stock.memory_v1.numpy().copy()
    mine = bank.copy()
    oracle.momentum_update(mine, v1.numpy(), y.numpy(), 0.5)
    pos = torch.from_numpy(bank)[y] * 0.5 + v1 * 0.5
    want = pos / pos.pow(2).sum(1, keepdim=True).pow(0.5)
    got = mine[y.numpy()]
    # canonical reduction order vs torch's: equal to within 2 ulp; untouched rows stay bit-identical
    assert np.max(np.abs(got - want.numpy()) / np.spacing(np.abs(want.numpy()))) <= 2
    mask = np.ones(len(bank), bool)
    mask[y.numpy()] = False
    assert np.array_equal(mine[mask], bank[mask])
    np.testing.assert_allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_momentum_update_duplicates_last_wins_and_shard(oracle):
    rng = np.random.default_rng(0)
    bank = rng.standard_normal((50, 64)).astype(np.float32)
    v = rng.standard_normal((4, 64)).astype(np.float32)
    y = np.array([7, 3, 7, 40], np.int64)
    a = bank.copy()
    oracle.momentum_update(a, v, y, 0.5)
    b = bank.copy()
    oracle.momentum_update(b, v[[2, 1, 3]], y[[2, 1, 3]], 0.5)   # only the last occurrence of 7
    assert np.array_equal(a, b)
    # shard [25,50): local row 15 == global row 40; rows of other shards are ignored
    sh = bank[25:].copy()
    oracle.momentum_update(sh, v, y, 0.5, row_begin=25, row_end=50)
    assert np.array_equal(sh[15], a[40]) and np.array_equal(np.delete(sh, 15, 0), np.delete(bank[25:], 15, 0))


def test_sharded_partials_sum_to_unsharded(oracle):
    stock, f_s, f_t, y, cidx = _setup(N=503)
    with torch.no_grad():
        v1 = stock.embed(f_s, stock.Ws, stock.bs).numpy()
        v2 = stock.embed(f_t, stock.Wt, stock.bt).numpy()
    b1, b2 = stock.memory_v1.numpy(), stock.memory_v2.numpy()
    full = oracle.crd_score(b1, b2, v1, v2, cidx.numpy(), 503, 0.07, 2000.0, 3000.0)
    acc = dict(loss_s=0.0, loss_t=0.0, sum_e1=0.0, count=0.0, grad_v1=0.0, out_v1=0.0)
    R = 4
    for r in range(R):
        lo, hi = 503 * r // R, 503 * (r + 1) // R
        part = oracle.crd_score(b1[lo:hi], b2[lo:hi], v1, v2, cidx.numpy(), 503, 0.07, 2000.0, 3000.0, row_begin=lo, row_end=hi)
        for k in acc:
            acc[k] = acc[k] + part[k]
    for k in acc:
        np.testing.assert_allclose(acc[k], full[k], rtol=1e-12)


# ---- edge cases: the shapes the kernels' parity tests lean on at their extremes ------------------------------------
def _plain_nce(bank1, bank2, v1, v2, idx, n_data, T, Z1, Z2, eps=1e-7):
    """The published loss in plain Python floats (one anchor at a time), float32 constants as in ContrastLoss."""
    B, K1 = idx.shape
    m = K1 - 1
    Pn = 1.0 / float(n_data)
    mPn = float(np.float32(m * Pn))          # P_neg.clone().fill_(m * Pn): a Python float stored into an fp32 tensor
    c = float(np.float32(m * Pn + eps))      # x.add(m * Pn + eps): the Python-float sum, rounded to fp32 once
    ls = lt = 0.0
    for b in range(B):
        for k in range(K1):
            r = int(idx[b, k])
            o1 = math.exp(float(np.dot(bank2[r].astype(np.float64), v1[b].astype(np.float64))) / T) / Z1
            o2 = math.exp(float(np.dot(bank1[r].astype(np.float64), v2[b].astype(np.float64))) / T) / Z2
            if k == 0:
                ls += math.log(o1 / (o1 + c))
                lt += math.log(o2 / (o2 + c))
            else:
                ls += math.log(mPn / (o1 + c))
                lt += math.log(mPn / (o2 + c))
    return -ls / B, -lt / B


@pytest.mark.parametrize("B,K", [(1, 7), (3, 0), (1, 0), (5, 1)])
def test_single_anchor_and_no_negatives(oracle, B, K):
    """B = 1 and K = 0 (only the positive column): the loss is the positive term alone and every gradient row is the
    positive's coefficient times its bank row."""
    rng = np.random.default_rng(11)
    N, D, T = 37, 8, 0.07
    b1 = oracle.l2_normalize(rng.standard_normal((N, D)).astype(np.float32))
    b2 = oracle.l2_normalize(rng.standard_normal((N, D)).astype(np.float32))
    v1 = oracle.l2_normalize(rng.standard_normal((B, D)).astype(np.float32))
    v2 = oracle.l2_normalize(rng.standard_normal((B, D)).astype(np.float32))
    idx = rng.integers(0, N, (B, K + 1))
    got = oracle.crd_score(b1, b2, v1, v2, idx, N, T, 50.0, 60.0)
    ws, wt = _plain_nce(b1, b2, v1, v2, idx, N, T, 50.0, 60.0)
    assert got["count"] == B * (K + 1)
    np.testing.assert_allclose([got["loss_s"], got["loss_t"]], [ws, wt], rtol=1e-12)
    if K == 0:
        c = float(np.float32(1e-7))
        for b in range(B):
            r = idx[b, 0]
            o1 = got["out_v1"][b, 0]
            coef = -c / (B * T * (o1 + c))
            np.testing.assert_allclose(got["grad_v1"][b], coef * b2[r].astype(np.float64), rtol=1e-9, atol=1e-300)


def test_empty_shard_and_boundary_rows(oracle):
    """A shard that owns none of the sampled rows contributes exactly nothing; rows on the shard's first and last index
    are counted by exactly one shard."""
    rng = np.random.default_rng(12)
    N, D, B, K, T = 64, 16, 4, 9, 0.07
    b1 = oracle.l2_normalize(rng.standard_normal((N, D)).astype(np.float32))
    b2 = oracle.l2_normalize(rng.standard_normal((N, D)).astype(np.float32))
    v1 = oracle.l2_normalize(rng.standard_normal((B, D)).astype(np.float32))
    v2 = oracle.l2_normalize(rng.standard_normal((B, D)).astype(np.float32))
    idx = rng.integers(0, 32, (B, K + 1))          # every sample lives in rows [0, 32)
    idx[0, 1], idx[1, 2], idx[2, 3], idx[3, 4] = 0, 15, 16, 31   # the boundaries of shards [0,16) and [16,32)
    empty = oracle.crd_score(b1[32:], b2[32:], v1, v2, idx, N, T, 40.0, 40.0, row_begin=32, row_end=64)
    assert empty["count"] == 0 and empty["loss_s"] == 0.0 and empty["loss_t"] == 0.0
    assert not empty["grad_v1"].any() and not empty["grad_v2"].any() and not empty["out_v1"].any()
    full = oracle.crd_score(b1, b2, v1, v2, idx, N, T, 40.0, 40.0)
    lo = oracle.crd_score(b1[:16], b2[:16], v1, v2, idx, N, T, 40.0, 40.0, row_begin=0, row_end=16)
    hi = oracle.crd_score(b1[16:32], b2[16:32], v1, v2, idx, N, T, 40.0, 40.0, row_begin=16, row_end=32)
    assert lo["count"] + hi["count"] == full["count"] == B * (K + 1)
    assert lo["count"] == int((idx < 16).sum()) and hi["count"] == int((idx >= 16).sum())
    np.testing.assert_allclose(lo["grad_v2"] + hi["grad_v2"], full["grad_v2"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(lo["loss_s"] + hi["loss_s"], full["loss_s"], rtol=1e-12)
    # the outputs of rows a shard does not own stay zero in that shard's out_v
    assert not lo["out_v1"][idx >= 16].any() and not hi["out_v1"][idx < 16].any()
