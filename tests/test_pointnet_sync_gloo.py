"""World-size-2 gloo test (CPU) of the host logic behind ShapeEncoderPC.sync_batchnorm(): the hand-off between two phases
sums exactly the accumulator blocks that crdpn_pointnet_sync_blocks names (right buffer, offset, count, dtype), in place,
and touches nothing else.  The kernels themselves need a GPU (tests/test_pointnet_sync_gpu.py)."""
import ctypes
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import __graft_entry__ as ge
        pkg = ge.load_package()
        from crdpn_b200 import pointnet as pn
        lib = pkg._native.lib()
        B, P, F = 6, 700, 256
        nb = ctypes.c_int(0)
        buf, off, cnt, f64 = (ctypes.c_int * 4)(), (ctypes.c_size_t * 4)(), (ctypes.c_int64 * 4)(), (ctypes.c_int * 4)()
        n_ctx, n_ws = ctypes.c_size_t(0), ctypes.c_size_t(0)
        assert lib.crdpn_pointnet_train_ctx_bytes(B, P, F, ctypes.byref(n_ctx)) == 0
        # the backward workspace size needs a device query; the blocks' offsets do not depend on the grid, so any buffer
        # that covers the largest block end is enough here
        size_ws = 0
        for sp in (3, 4, 5):
            assert lib.crdpn_pointnet_sync_blocks(B, P, F, sp, ctypes.byref(nb), buf, off, cnt, f64) == 0
            size_ws = max([size_ws] + [off[i] + cnt[i] * (8 if f64[i] else 4) for i in range(nb.value) if buf[i] == 1])
        for sp in range(6):
            assert lib.crdpn_pointnet_sync_blocks(B, P, F, sp, ctypes.byref(nb), buf, off, cnt, f64) == 0
            # buffers: 0 = train ctx (only its statistics head is needed), 1 = backward workspace, 2 / 3 = d_bn3_w / d_bn3_b
            owners = {0: torch.zeros(1 << 16, dtype=torch.uint8), 1: torch.zeros(size_ws + 64, dtype=torch.uint8),
                      2: torch.zeros(F, dtype=torch.float32), 3: torch.zeros(F, dtype=torch.float32)}
            g = torch.Generator().manual_seed(100 * sp + rank)
            for k in (0, 1):
                owners[k].copy_(torch.randint(0, 255, owners[k].shape, generator=g, dtype=torch.uint8))
            blocks = [(buf[i], off[i], cnt[i], f64[i]) for i in range(nb.value)]
            views = {}
            for (bf, o, c, d) in blocks:     # well-formed numbers inside the blocks (random bytes could be NaN patterns)
                dt = torch.float64 if d else torch.float32
                if bf in (0, 1):
                    v = owners[bf][o:o + c * (8 if d else 4)].view(dt)
                else:
                    v = owners[bf]
                v.copy_(torch.arange(c, dtype=dt) * (rank + 1) + 0.5 * bf)
                views[(bf, o)] = v
            before = {k: t.clone() for k, t in owners.items()}
            pn._sum_over_ranks(lib, (B, P, F), sp, None, {k: (t, t.data_ptr()) for k, t in owners.items()})
            for (bf, o, c, d) in blocks:
                dt = torch.float64 if d else torch.float32
                want = torch.arange(c, dtype=dt) * sum(r + 1 for r in range(world)) + 0.5 * bf * world
                assert torch.equal(views[(bf, o)], want), (sp, bf, o)
            for k in (0, 1):                 # everything outside the blocks is untouched
                mask = torch.ones(owners[k].numel(), dtype=torch.bool)
                for (bf, o, c, d) in blocks:
                    if bf == k:
                        mask[o:o + c * (8 if d else 4)] = False
                assert torch.equal(owners[k][mask], before[k][mask]), (sp, k)
        dist.barrier()
        q.put((rank, "ok"))
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def test_sum_over_ranks_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
