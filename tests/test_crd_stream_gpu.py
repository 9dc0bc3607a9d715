"""GPU parity of the EXPERIMENTAL bank-STREAMING formulation of the CRD step (variant 0x200, csrc/crd_stream.cuh; opt-in,
not the default path) against the CPU oracle
and against the gather formulation: same loss / gradients to fp32 rounding (1e-4 relative, north_star), bank rows updated
bit-identically (the momentum update is the same code), for interleaved and dense banks, shards, duplicates, ragged tile
counts and tiles nobody sampled."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REL32 = 1e-4


def _rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _mem(pkg, cuda, N, K, B, interleave=True, row_begin=0, row_end=None, seed=3):
    torch.manual_seed(seed)
    mem = pkg.ContrastMemory(128, N, K, 0.07, 0.5, interleave=interleave, row_begin=row_begin, row_end=row_end).to(cuda)
    g = torch.Generator().manual_seed(seed + 1)
    v1 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    y = torch.randperm(N, generator=g)[:B].to(cuda)
    cidx = torch.randint(0, N, (B, K + 1), generator=g).to(cuda)
    cidx[:, 0] = y
    return mem, v1, v2, y, cidx


@pytest.mark.parametrize("N,K,B,interleave", [(4096, 1023, 46, True), (4099, 2048, 48, True), (4100, 777, 5, False),
                                              (40000, 4096, 17, True), (1055, 300, 1, True)])
def test_stream_step_matches_oracle_and_gather(pkg, oracle, cuda, N, K, B, interleave):
    mem, v1, v2, y, cidx = _mem(pkg, cuda, N, K, B, interleave)
    if N == 40000:
        cidx[:, 1:] = cidx[:, 1:] % 20000   # the upper half of the bank is never sampled: tiles without records
        cidx[:, 0] = y
    b1 = mem.memory_v1.cpu().numpy().copy(); b2 = mem.memory_v2.cpu().numpy().copy()
    mem._freeze_z(v1, v2, cidx)
    hp = mem._host_params()
    want = oracle.crd_score(b1, b2, v1.cpu().numpy(), v2.cpu().numpy(), cidx.cpu().numpy(), N, 0.07, hp.Z1, hp.Z2)
    banks0 = (mem.memory_v1.clone(), mem.memory_v2.clone())
    # gather formulation
    mem.streaming = False
    res_g, g1_g, g2_g = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    res_g, g1_g, g2_g = res_g.clone(), g1_g.clone(), g2_g.clone()
    after_g = (mem.memory_v1.clone(), mem.memory_v2.clone())
    with torch.no_grad():
        mem.memory_v1.copy_(banks0[0]); mem.memory_v2.copy_(banks0[1])
    # streaming formulation
    mem.streaming = True
    res_s, g1_s, g2_s = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    torch.cuda.synchronize()
    assert _rel(res_s[0].item(), want["loss_s"]) < REL32 and _rel(res_s[1].item(), want["loss_t"]) < REL32
    assert _rel(g1_s.cpu(), want["grad_v1"]) < REL32 and _rel(g2_s.cpu(), want["grad_v2"]) < REL32
    assert _rel(res_s[5].item(), res_g[5].item()) < 1e-5
    assert _rel(g1_s, g1_g) < 1e-5 and _rel(g2_s, g2_g) < 1e-5
    assert res_s[4].item() == B * (K + 1)
    assert abs(res_s.view(torch.float32)[12].item() - res_s[5].item()) <= 1e-6 * abs(res_s[5].item())
    # the momentum update is the same code on the same inputs: bit-identical rows
    assert torch.equal(mem.memory_v1, after_g[0]) and torch.equal(mem.memory_v2, after_g[1])


def test_stream_step_on_a_shard_with_duplicates(pkg, oracle, cuda):
    N, K, B = 9000, 1500, 24
    lo, hi = 3000, 7003
    mem, v1, v2, y, cidx = _mem(pkg, cuda, N, K, B, True, lo, hi)
    y[B // 2:] = y[:B // 2]          # duplicate positives (tripled batches of the KD loop)
    cidx[:, 0] = y
    loc1, loc2 = mem.memory_v1.cpu().numpy().copy(), mem.memory_v2.cpu().numpy().copy()   # local row 0 = global row lo
    Z1, Z2 = 1234.5, 987.6
    want = oracle.crd_score(loc1, loc2, v1.cpu().numpy(), v2.cpu().numpy(), cidx.cpu().numpy(), N, 0.07, Z1, Z2,
                            row_begin=lo, row_end=hi)
    mem.streaming = True
    res, g1, g2 = mem._step(v1, v2, y, cidx, Z1, Z2)
    assert _rel(res[0].item(), want["loss_s"]) < REL32 and _rel(res[1].item(), want["loss_t"]) < REL32
    assert _rel(g1.cpu(), want["grad_v1"]) < REL32 and _rel(g2.cpu(), want["grad_v2"]) < REL32
    inshard = ((cidx >= lo) & (cidx < hi)).sum().item()
    assert res[4].item() == inshard


def test_crdloss_end_to_end_with_streaming(pkg, oracle, cuda):
    """CRDLoss(...).backward() through the two-call path with the streaming step == the gather step."""
    opt = type("Opt", (), dict(s_dim=64, t_dim=48, feat_dim=128, n_data=6000, nce_k=4096, nce_t=0.07, nce_m=0.5))()
    torch.manual_seed(5)
    a = pkg.CRDLoss(opt).to(cuda)
    b = pkg.CRDLoss(opt).to(cuda)
    b.load_state_dict(a.state_dict())
    a.contrast.streaming, b.contrast.streaming = False, True
    g = torch.Generator().manual_seed(6)
    for step in range(2):
        f_s, f_t = torch.randn(46, 64, generator=g).to(cuda), torch.randn(46, 48, generator=g).to(cuda)
        y = torch.randperm(6000, generator=g)[:46].to(cuda)
        cidx = torch.randint(0, 6000, (46, 4097), generator=g).to(cuda)
        cidx[:, 0] = y
        fa, fb = f_s.clone().requires_grad_(), f_s.clone().requires_grad_()
        a.zero_grad(); b.zero_grad()
        la = a(fa, f_t, y, cidx); la.backward()
        lb = b(fb, f_t, y, cidx); lb.backward()
        assert b.contrast._step_variant(46, 4097, 128) & 0x200
        assert _rel(lb.item(), la.item()) < 1e-5 and _rel(fb.grad, fa.grad) < 1e-4
        assert _rel(b.embed_s.linear.weight.grad, a.embed_s.linear.weight.grad) < 1e-4
        assert torch.equal(a.contrast.memory_v1, b.contrast.memory_v1)


def test_streaming_refuses_what_it_cannot_do(pkg, cuda):
    mem, v1, v2, y, cidx = _mem(pkg, cuda, 4096, 255, 49, True)
    mem._freeze_z(v1, v2, cidx)
    mem.streaming = True
    with pytest.raises(RuntimeError, match="batch <= 48"):
        mem._step(v1, v2, y, cidx, 10.0, 10.0)
    mem.streaming = False
    assert not (mem._step_variant(49, 256, 128) & 0x200)


# ---- bf16 banks: the tensor-core formulation (csrc/crd_tc_stream.cuh) -------------------------------------------------
RELBF = 1e-2   # north_star's bf16 tolerance


@pytest.mark.parametrize("N,K,B,interleave", [(4096, 1023, 46, True), (4099, 2048, 48, True), (4100, 777, 5, False),
                                              (40000, 4096, 17, True), (1055, 300, 1, True), (63, 200, 3, True),
                                              (20000, 16384, 46, True), (200000, 8192, 46, True)])
def test_tensor_core_stream_step_bf16_banks(pkg, oracle, cuda, N, K, B, interleave):
    """bf16 banks + streaming = the tcgen05 kernel: loss / gradients within the bf16 tolerance of the oracle evaluated on
    the same (bf16-valued) banks, close to the bf16 gather kernel, sample count exact, updated rows bit-identical."""
    torch.manual_seed(3)
    mem = pkg.ContrastMemory(128, N, K, 0.07, 0.5, interleave=interleave, bank_dtype=torch.bfloat16).to(cuda)
    g = torch.Generator().manual_seed(4)
    v1 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    y = torch.randperm(N, generator=g)[:B].to(cuda)
    cidx = torch.randint(0, N, (B, K + 1), generator=g).to(cuda)
    if N == 40000:
        cidx[:, 1:] = cidx[:, 1:] % 20000
    if N >= 20000:   # repeats of one row per anchor: the crowded-tile path (N = 20000) and the fast path's claim bits (N = 200000)
        cidx[:, 1:200] = cidx[:, 1:2]
        cidx[:, 300:340] = cidx[:, 300:301] ^ 1   # ... and of its pair-mate row (same 32-bit word of the operand image)
    cidx[:, 0] = y
    b1 = mem.memory_v1.float().cpu().numpy().copy(); b2 = mem.memory_v2.float().cpu().numpy().copy()
    mem._freeze_z(v1, v2, cidx)
    hp = mem._host_params()
    want = oracle.crd_score(b1, b2, v1.cpu().numpy(), v2.cpu().numpy(), cidx.cpu().numpy(), N, 0.07, hp.Z1, hp.Z2)
    banks0 = (mem.memory_v1.clone(), mem.memory_v2.clone())
    mem.streaming = False
    res_g, g1_g, g2_g = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    res_g, g1_g, g2_g = res_g.clone(), g1_g.clone(), g2_g.clone()
    after_g = (mem.memory_v1.clone(), mem.memory_v2.clone())
    with torch.no_grad():
        mem.memory_v1.copy_(banks0[0]); mem.memory_v2.copy_(banks0[1])
    mem.streaming = True
    assert mem._step_variant(B, K + 1, 128) & 0x200
    res_s, g1_s, g2_s = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    torch.cuda.synchronize()
    assert res_s[4].item() == B * (K + 1)
    assert _rel(res_s[0].item(), want["loss_s"]) < RELBF and _rel(res_s[1].item(), want["loss_t"]) < RELBF
    assert _rel(g1_s.cpu(), want["grad_v1"]) < RELBF and _rel(g2_s.cpu(), want["grad_v2"]) < RELBF
    assert _rel(res_s[5].item(), res_g[5].item()) < RELBF
    assert _rel(g1_s, g1_g) < RELBF and _rel(g2_s, g2_g) < RELBF
    assert torch.equal(mem.memory_v1, after_g[0]) and torch.equal(mem.memory_v2, after_g[1])
    # second step on the updated banks: the TMEM accumulators and barriers start clean every launch
    res_2, g1_2, _ = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    mem.streaming = False
    with torch.no_grad():
        mem.memory_v1.copy_(after_g[0]); mem.memory_v2.copy_(after_g[1])
    res_3, g1_3, _ = mem._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
    assert _rel(res_2[5].item(), res_3[5].item()) < RELBF and _rel(g1_2, g1_3) < RELBF


def test_tensor_core_stream_on_a_shard(pkg, oracle, cuda):
    N, K, B = 9000, 1500, 24
    lo, hi = 3000, 7003
    torch.manual_seed(3)
    mem = pkg.ContrastMemory(128, N, K, 0.07, 0.5, row_begin=lo, row_end=hi, bank_dtype=torch.bfloat16).to(cuda)
    g = torch.Generator().manual_seed(4)
    v1 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=1).to(cuda)
    y = torch.randperm(N, generator=g)[:B].to(cuda)
    y[B // 2:] = y[:B // 2]
    cidx = torch.randint(0, N, (B, K + 1), generator=g).to(cuda)
    cidx[:, 0] = y
    loc1, loc2 = mem.memory_v1.float().cpu().numpy().copy(), mem.memory_v2.float().cpu().numpy().copy()
    Z1, Z2 = 1234.5, 987.6
    want = oracle.crd_score(loc1, loc2, v1.cpu().numpy(), v2.cpu().numpy(), cidx.cpu().numpy(), N, 0.07, Z1, Z2,
                            row_begin=lo, row_end=hi)
    mem.streaming = True
    res, g1, g2 = mem._step(v1, v2, y, cidx, Z1, Z2)
    assert _rel(res[0].item(), want["loss_s"]) < RELBF and _rel(res[1].item(), want["loss_t"]) < RELBF
    assert _rel(g1.cpu(), want["grad_v1"]) < RELBF and _rel(g2.cpu(), want["grad_v2"]) < RELBF
    assert res[4].item() == ((cidx >= lo) & (cidx < hi)).sum().item()


def test_streaming_is_automatic_for_bf16_banks_only(pkg, cuda):
    """streaming=None (the default): bf16 banks stream when the step draws >= 2 samples per resident row; fp32 banks never."""
    bf = pkg.ContrastMemory(128, 8192, 4096, 0.07, 0.5, bank_dtype=torch.bfloat16).to(cuda)
    assert bf.streaming is None
    assert bf._step_variant(46, 4097, 128) & 0x200            # 188 K samples over 8 K rows
    assert not (bf._step_variant(2, 4097, 128) & 0x200)       # 8 K samples: gathering is cheaper than streaming the bank
    assert not (bf._step_variant(49, 4097, 128) & 0x200)      # unsupported batch: falls back silently in automatic mode
    bf.streaming = False
    assert not (bf._step_variant(46, 4097, 128) & 0x200)
    f32 = pkg.ContrastMemory(128, 8192, 4096, 0.07, 0.5).to(cuda)
    assert not (f32._step_variant(46, 4097, 128) & 0x200)


def test_crdloss_bf16_banks_default_path_is_the_tensor_core_step(pkg, cuda):
    """CRDLoss with bf16 banks, default settings: forward + backward run the streaming step and agree with the gather
    step within the bf16 tolerance; the banks receive bit-identical updates."""
    opt = type("Opt", (), dict(s_dim=64, t_dim=48, feat_dim=128, n_data=6000, nce_k=4096, nce_t=0.07, nce_m=0.5))()
    torch.manual_seed(5)
    a = pkg.CRDLoss(opt, bank_dtype=torch.bfloat16).to(cuda)
    b = pkg.CRDLoss(opt, bank_dtype=torch.bfloat16).to(cuda)
    b.load_state_dict(a.state_dict())
    a.contrast.streaming = False
    assert b.contrast.streaming is None and b.contrast._step_variant(46, 4097, 128) & 0x200
    g = torch.Generator().manual_seed(6)
    for step in range(2):
        f_s, f_t = torch.randn(46, 64, generator=g).to(cuda), torch.randn(46, 48, generator=g).to(cuda)
        y = torch.randperm(6000, generator=g)[:46].to(cuda)
        cidx = torch.randint(0, 6000, (46, 4097), generator=g).to(cuda)
        cidx[:, 0] = y
        fa, fb = f_s.clone().requires_grad_(), f_s.clone().requires_grad_()
        a.zero_grad(); b.zero_grad()
        la = a(fa, f_t, y, cidx); la.backward()
        lb = b(fb, f_t, y, cidx); lb.backward()
        assert _rel(lb.item(), la.item()) < RELBF and _rel(fb.grad, fa.grad) < RELBF
        assert _rel(b.embed_s.linear.weight.grad, a.embed_s.linear.weight.grad) < RELBF
        assert torch.equal(a.contrast.memory_v1, b.contrast.memory_v1)
