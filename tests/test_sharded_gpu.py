"""GPU tests of the sharded CRD path: world_size 1 (runs on the driver's single B200) must reproduce CRDLoss
bit for bit; world_size 2 over NCCL (needs 2 GPUs: `gpurun --gpus 2`) must match the unsharded module."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _opt(**kw):
    base = dict(s_dim=64, t_dim=48, feat_dim=128, n_data=6000, nce_k=2048, nce_t=0.07, nce_m=0.5)
    base.update(kw)
    return type("Opt", (), base)()


def _inputs(opt, B, seed=5):
    g = torch.Generator().manual_seed(seed)
    f_s, f_t = torch.randn(B, opt.s_dim, generator=g), torch.randn(B, opt.t_dim, generator=g)
    y = torch.randperm(opt.n_data, generator=g)[:B]
    cidx = torch.randint(0, opt.n_data, (B, opt.nce_k + 1), generator=g)
    cidx[:, 0] = y
    return f_s, f_t, y, cidx


def test_world1_sharded_equals_unsharded(pkg, cuda):
    """Same kernels, same order: gradients and bank rows are bit-identical; the loss goes through the packed fp32
    exchange buffer in the sharded module (fp64 -> fp32 once), so it agrees to fp32 rounding."""
    opt = _opt()
    torch.manual_seed(1)
    a = pkg.CRDLoss(opt).to(cuda)
    b = pkg.ShardedCRDLoss(opt, rank=0, world_size=1).to(cuda)
    b.load_state_dict(a.state_dict(), strict=False)
    f_s, f_t, y, cidx = [t.to(cuda) for t in _inputs(opt, 46)]
    for step in range(2):
        fa, fb = f_s.clone().requires_grad_(), f_s.clone().requires_grad_()
        la = a(fa, f_t, y, cidx); la.backward()
        lb = b(fb, f_t, y, cidx); lb.backward()
        assert abs(la.item() - lb.item()) <= 2e-7 * abs(la.item()) and torch.equal(fa.grad, fb.grad)
        assert torch.equal(a.contrast.memory_v1, b.contrast.memory_v1)
        assert torch.equal(a.contrast.params, b.contrast.params)


def test_world1_peer_memory_step_equals_unsharded(pkg, cuda):
    """comm="p2p" with a single rank: the exchange kernels talk to this rank's own buffer, so the whole sharded call
    chain (crdpn_crd_loss_forward_sharded / crdpn_crd_step_sharded: all-gather kernel, scoring pass, reduction kernel with
    the sum over ranks fused in) runs on one GPU and must reproduce the unsharded step: gradients and bank rows bit for
    bit (one rank: no re-association), loss to fp32 rounding."""
    opt = _opt()
    torch.manual_seed(1)
    a = pkg.CRDLoss(opt).to(cuda)
    b = pkg.ShardedCRDLoss(opt, rank=0, world_size=1, comm="p2p").to(cuda)
    b.load_state_dict(a.state_dict(), strict=False)
    f_s, f_t, y, cidx = [t.to(cuda) for t in _inputs(opt, 46)]
    for step in range(3):   # step 0 freezes Z through the general path, steps 1-2 take the one-call path
        fa, fb = f_s.clone().requires_grad_(), f_s.clone().requires_grad_()
        la = a(fa, f_t, y, cidx); la.backward()
        lb = b(fb, f_t, y, cidx); lb.backward()
        assert abs(la.item() - lb.item()) <= 2e-7 * abs(la.item())
        assert torch.equal(fa.grad, fb.grad)
        assert torch.equal(a.contrast.memory_v1, b.contrast.memory_v1) and torch.equal(a.contrast.memory_v2, b.contrast.memory_v2)
    # the device-resident entry point (what bench.py times at N > 1), twice in a row (same-kind exchanges back to back)
    with torch.no_grad():
        v1, v2 = a.embed_s(f_s).contiguous(), a.embed_t(f_t).contiguous()
    hp = a.contrast._host_params()
    out = None
    for _ in range(2):
        res, g1, g2 = a.contrast._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
        out = b.contrast.step_resident(v1, v2, y, cidx, out)
        B, D = v1.shape
        assert torch.equal(out["reduced"][:B * D].view(B, D), g1) and torch.equal(out["reduced"][B * D:2 * B * D].view(B, D), g2)
        want = (res[0] + res[1]).item()
        assert abs(out["reduced"][2 * B * D + 5].item() - want) <= 2e-7 * abs(want)
        assert torch.equal(a.contrast.memory_v1, b.contrast.memory_v1)


def _shard_module(pkg, cuda, N, K, lo, hi, variant, seed=3):
    """One rank's ContrastMemory of a bank sharded three ways, on this GPU alone (exchange kernels talk to themselves)."""
    m = pkg.ShardedContrastMemory(128, N, K, rank=0, world_size=1, comm="p2p", seed=5).to(cuda)
    m.row_begin, m.row_end = lo, hi
    g = torch.Generator().manual_seed(seed)
    m.memory_v1 = torch.nn.functional.normalize(torch.randn(hi - lo, 128, generator=g)).to(cuda)
    m.memory_v2 = torch.nn.functional.normalize(torch.randn(hi - lo, 128, generator=g)).to(cuda)
    m._relayout()
    with torch.no_grad():
        m.params[2], m.params[3] = 1.0e5, 1.1e5
    m._host = None
    m.variant = variant
    return m


@pytest.mark.parametrize("shape", [dict(B=46, K=2048, N=6000), dict(B=7, K=5000, N=3001), dict(B=64, K=300, N=50000)])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_compact_prepass_matches_the_scanning_kernel(pkg, oracle, cuda, shape, dtype):
    """The filter pre-pass of the row-sharded step (crd_shard_filter_kernel + compact scoring; variant bit 6 forces it on,
    bit 5 off) against the scanning kernel and the CPU oracle on a shard that owns a THIRD of the rows: same loss /
    gradients up to fp32 summation order, updated rows bit-identical, the same bits run to run; int64 and int32 lists."""
    B, K, N = shape["B"], shape["K"], shape["N"]
    lo, hi = N // 3, 2 * N // 3
    g = torch.Generator().manual_seed(B + K)
    v1 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g)).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, 128, generator=g)).to(cuda)
    y = torch.randperm(N, generator=g)[:B].to(cuda)
    y[0] = lo + 1                                      # at least one positive inside the shard
    cidx = torch.randint(0, N, (B, K + 1), generator=g).to(cuda)
    cidx[:, 0] = y
    outs = []
    for variant in (0x20, 0x40, 0x40):   # scanning kernel, pre-pass, pre-pass again
        m = _shard_module(pkg, cuda, N, K, lo, hi, variant)
        if not outs:
            b1, b2 = m.memory_v1.cpu().numpy().copy(), m.memory_v2.cpu().numpy().copy()
        out = m.step_resident(v1, v2, y, cidx.to(dtype))
        torch.cuda.synchronize()
        outs.append((out["reduced"].clone(), m.memory_v1.clone(), m.memory_v2.clone()))
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
    n = 2 * B * 128
    assert rel(outs[1][0][:n], outs[0][0][:n]) < 1e-5
    assert abs(outs[1][0][n + 5].item() - outs[0][0][n + 5].item()) <= 1e-5 * abs(outs[0][0][n + 5].item())
    assert torch.equal(outs[1][1], outs[0][1]) and torch.equal(outs[1][2], outs[0][2])
    assert torch.equal(outs[1][0][:n + 6], outs[2][0][:n + 6])   # deterministic: identical bits run to run (words 6, 7 unused)
    r = oracle.crd_score(b1, b2, v1.cpu().numpy(), v2.cpu().numpy(), cidx.cpu().numpy(), N, 0.07, 1.0e5, 1.1e5,
                         row_begin=lo, row_end=hi, want_out=False)
    assert rel(outs[1][0][:B * 128].view(B, 128).cpu(), torch.from_numpy(r["grad_v1"])) < 1e-4
    assert rel(outs[1][0][B * 128:n].view(B, 128).cpu(), torch.from_numpy(r["grad_v2"])) < 1e-4
    assert abs(outs[1][0][n + 5].item() - (r["loss_s"] + r["loss_t"])) <= 1e-4 * abs(r["loss_s"] + r["loss_t"])


@pytest.mark.parametrize("variant", [0x20, 0x40])
def test_in_shard_negatives_drawn_inside_the_sharded_step(pkg, oracle, cuda, variant):
    """ShardedCRDLoss(comm="p2p", local_negatives=True), second call: one foreign call in which the scoring pass (bit 5) or
    the filter pre-pass (bit 6) draws the in-shard negatives itself -- against the oracle's draw + scorer."""
    import numpy as np
    opt = _opt(nce_k=700)
    torch.manual_seed(2)
    m = pkg.ShardedCRDLoss(opt, rank=0, world_size=1, local_negatives=True, comm="p2p", seed=77).to(cuda)
    m.contrast.variant = variant
    f_s, f_t, y, _ = [t.to(cuda) for t in _inputs(opt, 9)]
    m(f_s, f_t, y, None)                                  # first call: freezes Z (general path), updates the banks
    b1 = m.contrast.memory_v1.cpu().numpy().copy(); b2 = m.contrast.memory_v2.cpu().numpy().copy()
    loss = m(f_s, f_t, y, None)
    prob, alias = oracle.alias_build(np.ones(opt.n_data, np.float32))
    cidx = oracle.alias_draw_contrast(prob, alias, y.cpu().numpy(), 701, seed=77 + 7919, offset=9 * 701)
    with torch.no_grad():
        v1 = m.embed_s(f_s).cpu().numpy(); v2 = m.embed_t(f_t).cpu().numpy()
    Z1, Z2 = m.contrast.params[2].item(), m.contrast.params[3].item()
    want = oracle.crd_score(b1, b2, v1, v2, cidx, opt.n_data, 0.07, Z1, Z2)
    assert abs(loss.item() - (want["loss_s"] + want["loss_t"])) < 1e-4 * abs(want["loss_s"] + want["loss_t"])


def test_local_negatives_world1_matches_oracle(pkg, oracle, cuda):
    """local_negatives with internal sampling: indices are in-shard draws of the rank's own Philox stream."""
    opt = _opt(nce_k=512)
    torch.manual_seed(2)
    m = pkg.ShardedCRDLoss(opt, rank=0, world_size=1, local_negatives=True, seed=77).to(cuda)
    f_s, f_t, y, _ = [t.to(cuda) for t in _inputs(opt, 8)]
    b1 = m.contrast.memory_v1.cpu().numpy().copy(); b2 = m.contrast.memory_v2.cpu().numpy().copy()
    loss = m(f_s, f_t, y, None)
    import numpy as np
    prob, alias = oracle.alias_build(np.ones(opt.n_data, np.float32))
    cidx = oracle.alias_draw_contrast(prob, alias, y.cpu().numpy(), 513, seed=77 + 7919, offset=0)
    with torch.no_grad():
        v1 = m.embed_s(f_s).cpu().numpy(); v2 = m.embed_t(f_t).cpu().numpy()
    Z1, Z2 = m.contrast.params[2].item(), m.contrast.params[3].item()
    want = oracle.crd_score(b1, b2, v1, v2, cidx, opt.n_data, 0.07, Z1, Z2)
    assert abs(loss.item() - (want["loss_s"] + want["loss_t"])) < 1e-4 * abs(want["loss_s"] + want["loss_t"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q, comm="dist"):
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import __graft_entry__ as ge
        pkg = ge.load_package()
        opt = _opt(n_data=6001)
        torch.manual_seed(1)
        ref = pkg.CRDLoss(opt).to(dev)                       # the unsharded module, same weights on every rank
        sh = pkg.ShardedCRDLoss(opt, comm=comm).to(dev)
        lo, hi = sh.contrast.row_begin, sh.contrast.row_end
        with torch.no_grad():
            for n_ in ("embed_s", "embed_t"):
                getattr(sh, n_).load_state_dict(getattr(ref, n_).state_dict())
            sh.contrast.memory_v1.copy_(ref.contrast.memory_v1[lo:hi]); sh.contrast.memory_v2.copy_(ref.contrast.memory_v2[lo:hi])
        B = 46
        f_s, f_t, y, cidx = [t.to(dev) for t in _inputs(opt, B)]
        counts = [B * (r + 1) // world - B * r // world for r in range(world)]
        a0 = sum(counts[:rank]); sl = slice(a0, a0 + counts[rank])
        rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
        for step in range(2):
            fr = f_s.clone().requires_grad_()
            lr = ref(fr, f_t, y, cidx); lr.backward()
            fl = f_s[sl].clone().requires_grad_()
            ls = sh(fl, f_t[sl], y[sl], cidx); ls.backward()
            assert rel(ls, lr) < 1e-5, (ls.item(), lr.item())
            assert rel(fl.grad, fr.grad[sl]) < 1e-4
            assert rel(sh.contrast.params[2:4], ref.contrast.params[2:4]) < 1e-5
            # owner-only momentum update: shard rows are BIT-identical to the unsharded bank's rows
            assert torch.equal(sh.contrast.memory_v1, ref.contrast.memory_v1[lo:hi])
            assert torch.equal(sh.contrast.memory_v2, ref.contrast.memory_v2[lo:hi])
        if comm == "p2p":   # the device-resident entry point bench.py times: crdpn_crd_step_sharded
            with torch.no_grad():
                v1, v2 = ref.embed_s(f_s).contiguous(), ref.embed_t(f_t).contiguous()
            hp = ref.contrast._host_params()
            out = None
            for _ in range(3):
                res, g1, g2 = ref.contrast._step(v1, v2, y, cidx, hp.Z1, hp.Z2)
                out = sh.contrast.step_resident(v1[sl].contiguous(), v2[sl].contiguous(), y[sl].contiguous(), cidx, out)
                red = out["reduced"]
                assert rel(red[:B * 128].view(B, 128), g1) < 1e-4 and rel(red[B * 128:2 * B * 128].view(B, 128), g2) < 1e-4
                assert rel(red[2 * B * 128 + 5], res[0] + res[1]) < 1e-5
                assert torch.equal(sh.contrast.memory_v1, ref.contrast.memory_v1[lo:hi])
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("comm", ["dist", "p2p"])
def test_world2_nccl_matches_unsharded(pkg, comm):
    """comm="dist": NCCL collectives; comm="p2p": the two exchanges as single kernels over NVLink peer memory."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q, comm)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_world1_graphed_sharded_step_equals_the_eager_loop(pkg, cuda):
    """The sharded forward + backward (peer-memory all-gather, in-shard negatives drawn inside the scoring pass, reduction with
    the sum over ranks fused in) captured in a CUDA graph by GraphedStep: replays must reproduce the eager loop bit for bit,
    with fresh negatives on every replay (device-resident sampler offset) -- on one rank, so it runs on the driver's GPU."""
    opt = _opt(nce_k=1024)
    B = 16
    mods = []
    for _ in range(2):
        torch.manual_seed(3)
        mods.append(pkg.ShardedCRDLoss(opt, rank=0, world_size=1, comm="p2p", local_negatives=True, fixed_local_batch=True,
                                       seed=21).to(cuda))
    a, b = mods
    b.load_state_dict(a.state_dict(), strict=False)
    gen = torch.Generator().manual_seed(8)
    batches = [(torch.randn(B, opt.s_dim, generator=gen).pin_memory(), torch.randn(B, opt.t_dim, generator=gen).pin_memory(),
                torch.randperm(opt.n_data, generator=gen)[:B].pin_memory()) for _ in range(4)]

    def eager(m, batch):
        f_s = batch[0].to(cuda).requires_grad_(True)
        m.zero_grad(set_to_none=True)
        loss = m(f_s, batch[1].to(cuda), batch[2].to(cuda))
        loss.backward()
        return loss, f_s

    for m in (a, b):          # step 0 freezes Z through the general path, step 1 takes the one-call path
        eager(m, batches[0])
        eager(m, batches[0])
    b.contrast.device_sampler_offset()

    def fwd_bwd(f_s, f_t, idx):
        loss = b(f_s, f_t, idx)
        loss.backward()
        return loss

    warm = 2
    step = pkg.GraphedStep(fwd_bwd, batches[0], cuda, grad_inputs=(0,), zero_grad=lambda: b.zero_grad(set_to_none=True), warmup=warm)
    for _ in range(warm):
        eager(a, batches[0])
    step.stage(*batches[1])
    for i in range(1, 4):
        step.run()
        if i + 1 < 4:
            step.stage(*batches[i + 1])
        want, f_s = eager(a, batches[i])
        assert step.collect() == want.item(), i
        assert torch.equal(step.static[0].grad, f_s.grad)
        assert torch.equal(b.embed_s.linear.weight.grad, a.embed_s.linear.weight.grad)
    assert torch.equal(b.contrast.memory_v1, a.contrast.memory_v1) and torch.equal(b.contrast.memory_v2, a.contrast.memory_v2)
