"""Rank-synchronised (SyncBN-style) PointNet training: ShapeEncoderPC.sync_batchnorm() runs the train forward / backward
in phases with a sum over ranks of the per-channel accumulators in between (crdpn_pointnet_*_phased,
crdpn_pointnet_sync_blocks).

world 1 (runs on the driver's single B200): the phased path must reproduce the one-call path BIT FOR BIT.
world 2 (needs `gpurun --gpus 2`): two ranks with half the clouds each must reproduce the single-GPU run over the whole
batch -- features, running statistics, and parameter gradients (which come out already summed over ranks) -- within the
bf16 recipe's tolerance (the tile partition differs, so a few bf16 roundings / near-tied arg-max points may flip)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _state(F, seed=46):
    g = torch.Generator().manual_seed(seed)
    st = {}
    for cin, cout, n in ((3, 64, 1), (64, 128, 2), (128, F, 3)):
        bound = 1.0 / (cin ** 0.5)
        st[f"conv{n}.weight"] = (torch.rand(cout, cin, 1, generator=g) * 2 - 1) * bound
        st[f"conv{n}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        st[f"bn{n}.weight"] = torch.randn(cout, generator=g)
        st[f"bn{n}.bias"] = torch.randn(cout, generator=g)
        st[f"bn{n}.running_mean"] = torch.randn(cout, generator=g) * 0.2
        st[f"bn{n}.running_var"] = torch.rand(cout, generator=g) * 1.5 + 0.5
        st[f"bn{n}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return st


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _worker(rank, world, port, q, comm="dist"):
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import __graft_entry__ as ge
        pkg = ge.load_package()
        F, B, P = 256, 6, 700
        g = torch.Generator().manual_seed(7)
        x = torch.rand(B, 3, P, generator=g).to(dev)
        gout = torch.randn(B, F, generator=g).to(dev)
        st = _state(F)
        ref = pkg.ShapeEncoderPC(F); ref.load_state_dict(st); ref = ref.to(dev).train()
        syn = pkg.ShapeEncoderPC(F); syn.load_state_dict(st); syn = syn.to(dev).train().sync_batchnorm(comm=comm, equal_batches=(comm == "p2p"))
        counts = [B * (r + 1) // world - B * r // world for r in range(world)]
        a0 = sum(counts[:rank]); sl = slice(a0, a0 + counts[rank])
        for step in range(2):
            for m in (ref, syn):
                m.zero_grad(set_to_none=True)
            out_ref = ref(x); out_ref.backward(gout)
            out_syn = syn(x[sl].contiguous()); out_syn.backward(gout[sl].contiguous())
            if world == 1:
                assert torch.equal(out_syn, out_ref)
                for (n1, p1), (_, p2) in zip(ref.named_parameters(), syn.named_parameters()):
                    assert torch.equal(p1.grad, p2.grad), n1
                for (n1, b1), (_, b2) in zip(ref.named_buffers(), syn.named_buffers()):
                    assert torch.equal(b1, b2), n1
            else:
                assert _rel(out_syn, out_ref[sl]) < 2e-3, _rel(out_syn, out_ref[sl])
                for (n1, b1), (_, b2) in zip(ref.named_buffers(), syn.named_buffers()):
                    if b1.is_floating_point():
                        assert _rel(b2, b1) < 1e-5, (n1, _rel(b2, b1))
                    else:
                        assert torch.equal(b1, b2), n1
                for (n1, p1), (_, p2) in zip(ref.named_parameters(), syn.named_parameters()):
                    if n1.startswith("conv") and n1.endswith("bias"):
                        assert p2.grad.abs().max() <= 1e-6 * max(1.0, p1.grad.abs().max().item()) + 1e-6, n1   # exactly-zero gradients
                        continue
                    assert _rel(p2.grad, p1.grad) < 1e-2, (n1, _rel(p2.grad, p1.grad))
                # identical on every rank: the gradients are functions of globally summed accumulators only
                flat = torch.cat([p.grad.reshape(-1) for p in syn.parameters()])
                other = flat.clone()
                dist.broadcast(other, src=0)
                assert torch.equal(flat, other)
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def _run(world, comm="dist"):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, comm)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


@pytest.mark.parametrize("comm", ["dist", "p2p"])
def test_world1_phased_equals_one_call_bitwise(pkg, comm):
    """comm="p2p": the hand-offs run through crdpn_p2p_allreduce_blocks against this rank's own exchange buffer (in-place
    sum of one rank = identity, bit for bit, float32 and float64 blocks alike)."""
    _run(1, comm)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("comm", ["dist", "p2p"])
def test_world2_sync_matches_single_gpu_full_batch(pkg, comm):
    _run(2, comm)


def test_sync_blocks_table(pkg):
    """crdpn_pointnet_sync_blocks: block counts, dtypes and sizes per sync point."""
    import ctypes
    lib = pkg._native.lib()
    nb = ctypes.c_int(0)
    buf, off, cnt, f64 = (ctypes.c_int * 4)(), (ctypes.c_size_t * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int * 4)()
    want = {0: [(0, 16, 1)], 1: [(0, 256, 1)], 2: [(0, 2048, 1)],
            3: [(1, 128, 1), (1, 1024 * 128, 0), (2, 1024, 0), (3, 1024, 0)], 4: [(1, 256, 1), (1, 128 * 64 + 2 * 128 * 128, 0)],
            5: [(1, 320, 1)]}
    for sp, blocks in want.items():
        assert lib.crdpn_pointnet_sync_blocks(160, 2500, 1024, sp, ctypes.byref(nb), buf, off, cnt, f64) == 0
        assert [(buf[i], cnt[i], f64[i]) for i in range(nb.value)] == blocks, sp
        assert all(off[i] % 8 == 0 for i in range(nb.value))
    assert lib.crdpn_pointnet_sync_blocks(160, 2500, 1024, 6, ctypes.byref(nb), buf, off, cnt, f64) != 0
