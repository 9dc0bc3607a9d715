"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, the packed all-gather of anchors,
the packed all-reduce of partials, first-call Z over ranks, owner-only bank updates and autograd routing.
The per-rank kernel is replaced by the CPU oracle (tests may use it); everything else is the product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_oracle_memory(pkg, oracle):
    class OracleShardedMemory(pkg.ShardedContrastMemory):
        """Host logic under test; the CUDA kernel calls swapped for the oracle on CPU tensors."""

        def _check_device(self, t, name):
            pass

        def _oracle(self, v1, v2, idx, Z1, Z2):
            hp = self._host_params()
            r = oracle.crd_score(self.memory_v1.numpy(), self.memory_v2.numpy(), v1.numpy(), v2.numpy(), idx.numpy(),
                                 self.nLem, hp.T, Z1, Z2, row_begin=self.row_begin, row_end=self.row_end,
                                 k_total=self.k_total)
            res = torch.tensor([r["loss_s"], r["loss_t"], r["sum_e1"], r["sum_e2"], r["count"], 0, 0, 0], dtype=torch.float64)
            return res, torch.from_numpy(r["grad_v1"]).float(), torch.from_numpy(r["grad_v2"]).float()

        def _score(self, v1, v2, idx, Z1, Z2, want_out=False, result=None):
            res, g1, g2 = self._oracle(v1, v2, idx, Z1, Z2)
            return res, g1, g2, None, None

        def _step(self, v1, v2, y, idx, Z1, Z2):
            res, g1, g2 = self._oracle(v1, v2, idx, Z1, Z2)
            for mem, v in ((self.memory_v1, v1), (self.memory_v2, v2)):
                oracle.momentum_update(mem.numpy(), v.numpy(), y.numpy(), self._host_params().m,
                                       row_begin=self.row_begin, row_end=self.row_end)
            return res, g1, g2

    return OracleShardedMemory


def _worker(rank, world, port, local_negatives, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import __graft_entry__ as ge
        from oracle import crd_oracle as oracle
        from oracle.crd_oracle import StockCRD
        pkg = ge.load_package()
        N, D, K, B = 301, 32, 40, 7
        opt = type("Opt", (), dict(s_dim=20, t_dim=12, feat_dim=D, n_data=N, nce_k=K, nce_t=0.07, nce_m=0.5))()
        stock = StockCRD(opt.s_dim, opt.t_dim, D, N, K * (world if local_negatives else 1), 0.07, 0.5, seed=3)
        crit = pkg.ShardedCRDLoss(opt, local_negatives=local_negatives, interleave=False)
        crit.contrast.__class__ = _make_oracle_memory(pkg, oracle)

        class HostEmbed(pkg.Embed):   # the product Embed is CUDA-only; the host logic under test needs a CPU stand-in
            def forward(self, x):
                return self.l2norm(self.linear(x.view(x.shape[0], -1)))

        crit.embed_s.__class__ = HostEmbed
        crit.embed_t.__class__ = HostEmbed
        lo, hi = pkg.shard_bounds(N, world, rank)
        assert (crit.contrast.row_begin, crit.contrast.row_end) == (lo, hi)
        with torch.no_grad():
            crit.embed_s.linear.weight.copy_(stock.Ws); crit.embed_s.linear.bias.copy_(stock.bs)
            crit.embed_t.linear.weight.copy_(stock.Wt); crit.embed_t.linear.bias.copy_(stock.bt)
            crit.contrast.memory_v1.copy_(stock.memory_v1[lo:hi]); crit.contrast.memory_v2.copy_(stock.memory_v2[lo:hi])
        g = torch.Generator().manual_seed(11)
        f_s, f_t = torch.randn(B, opt.s_dim, generator=g), torch.randn(B, opt.t_dim, generator=g)
        y = torch.randperm(N, generator=g)[:B]
        if local_negatives:  # every rank brings K negatives of its own shard; the union is what one GPU would see
            per_rank = [torch.randint(*pkg.shard_bounds(N, world, r), (B, K), generator=g) for r in range(world)]
            cidx_full = torch.cat([y.view(-1, 1)] + per_rank, dim=1)
            cidx = torch.cat([y.view(-1, 1), per_rank[rank]], dim=1)
        else:
            cidx_full = torch.randint(0, N, (B, K + 1), generator=g)
            cidx_full[:, 0] = y
            cidx = cidx_full
        f_s_all, f_t_all, y_all, cidx_all, cidx_full_all = f_s, f_t, y, cidx, cidx_full
        for step in range(3):
            # uneven data-parallel split of the 7 anchors; in the last step the batch shrinks ON RANK 1 ONLY (a ragged
            # final batch): the per-rank sizes must be re-exchanged by every rank, not only by the rank that changed
            counts = [4, 3] if step < 2 else [4, 2]
            nb = sum(counts)
            f_s, f_t, y, cidx, cidx_full = f_s_all[:nb], f_t_all[:nb], y_all[:nb], cidx_all[:nb], cidx_full_all[:nb]
            a0 = sum(counts[:rank])
            sl = slice(a0, a0 + counts[rank])
            fs_l = f_s[sl].clone().requires_grad_()
            ft_l = f_t[sl].clone().requires_grad_()
            crit.zero_grad()
            loss = crit(fs_l, ft_l, y[sl], cidx)
            loss.backward()
            crit.allreduce_embed_grads()
            fs_c, ft_c = f_s.clone().requires_grad_(), f_t.clone().requires_grad_()
            for p in (stock.Ws, stock.bs, stock.Wt, stock.bt):
                p.grad = None
            want = stock.loss(fs_c, ft_c, y, cidx_full)
            want.backward()
            rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()
            assert rel(loss, want) < 1e-4, (loss.item(), want.item())
            assert rel(torch.tensor(crit.contrast._host_params().Z1), torch.tensor(stock.Z1)) < 1e-4
            assert rel(fs_l.grad, fs_c.grad[sl]) < 1e-4 and rel(ft_l.grad, ft_c.grad[sl]) < 1e-4
            assert rel(crit.embed_s.linear.weight.grad, stock.Ws.grad) < 1e-4
            assert rel(crit.embed_t.linear.bias.grad, stock.bt.grad) < 1e-4
            # owner-only updates: my shard equals the slice of the single-process bank, other rows untouched there
            assert rel(crit.contrast.memory_v1, stock.memory_v1[lo:hi]) < 1e-5
            assert rel(crit.contrast.memory_v2, stock.memory_v2[lo:hi]) < 1e-5
            with torch.no_grad():  # re-sync rows so the second step starts from identical banks
                crit.contrast.memory_v1.copy_(stock.memory_v1[lo:hi]); crit.contrast.memory_v2.copy_(stock.memory_v2[lo:hi])
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("local_negatives", [False, True])
def test_sharded_crd_world2_matches_single_process(pkg, oracle, local_negatives):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, local_negatives, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_pack_unpack_roundtrip_uneven(pkg):
    from importlib import import_module
    sh = import_module("crdpn_b200.sharded")
    v1, v2 = torch.randn(3, 8), torch.randn(3, 8)
    y = torch.tensor([2**40 + 5, 7, 0])
    a = sh.pack_anchor_rows(v1, v2, y, 4)
    b = sh.pack_anchor_rows(v1[:2] + 1, v2[:2] + 1, y[:2] + 1, 4)
    g1, g2, gy = sh.unpack_anchor_rows(torch.cat([a, b]), [3, 2], 4, 8)
    assert torch.equal(g1, torch.cat([v1, v1[:2] + 1])) and torch.equal(g2, torch.cat([v2, v2[:2] + 1]))
    assert torch.equal(gy, torch.cat([y, y[:2] + 1]))
    assert [sh.shard_bounds(10, 4, r) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
