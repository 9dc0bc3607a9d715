"""Out-of-bounds guard for the C ABI entry points added in round 1 (compute-sanitizer is not available on this pool): every
output / workspace buffer handed to the library sits between two 4 KB guard bands filled with a pattern; after the call
the bands must be untouched and the payload must have been written."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096
PAT = 0xA5


class Guarded:
    """`nbytes` of device memory (256-byte aligned) between two guard bands."""

    def __init__(self, nbytes, dev, fill=0xEE):
        self.n = (int(nbytes) + 255) // 256 * 256
        self.buf = torch.full((self.n + 2 * GUARD,), PAT, dtype=torch.uint8, device=dev)
        self.buf[GUARD:GUARD + self.n] = fill
        self.ptr = self.buf.data_ptr() + GUARD
        assert self.ptr % 256 == 0
        self.nbytes = int(nbytes)
        self.fill = fill

    def view(self, dtype, count):
        return self.buf[GUARD:GUARD + self.n].view(dtype)[:count]

    def intact(self):
        lo, hi = self.buf[:GUARD], self.buf[GUARD + self.n:]
        pad = self.buf[GUARD + self.nbytes:GUARD + self.n]   # alignment slack after the payload: must stay as filled
        return bool((lo == PAT).all() and (hi == PAT).all() and (pad == self.fill).all())


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def test_nce_kd_guards(pkg, cuda):
    lib = pkg._native.lib()
    for B, C, weighting, p in ((46, 200, 0, 0.3), (7, 33, 5, 0.0), (138, 200, 1, 0.0)):
        g = torch.Generator().manual_seed(B)
        a, q = torch.randn(B, C, generator=g).to(cuda), torch.randn(B, C, generator=g).to(cuda)
        lab = (torch.rand(B, 3, generator=g) * 180).to(cuda)
        n = ctypes.c_size_t(0)
        assert lib.crdpn_nce_kd_workspace_bytes(B, C, ctypes.byref(n)) == 0
        ws, loss = Guarded(n.value, cuda), Guarded(4, cuda)
        d_ori, d_pos = Guarded(B * C * 4, cuda), Guarded(B * C * 4, cuda)
        rc = lib.crdpn_nce_kd_forward(a.data_ptr(), q.data_ptr(), lab.data_ptr(), B, C, 0.5, weighting, p, 11, 3, loss.ptr, ws.ptr,
                                      n.value, _st(cuda))
        assert rc == 0
        rc = lib.crdpn_nce_kd_backward(None, B, C, 0.5, p, 11, 3, ws.ptr, n.value, d_ori.ptr, d_pos.ptr, _st(cuda))
        assert rc == 0
        torch.cuda.synchronize()
        assert all(x.intact() for x in (ws, loss, d_ori, d_pos))
        assert torch.isfinite(loss.view(torch.float32, 1)).all() and torch.isfinite(d_ori.view(torch.float32, B * C)).all()
        assert torch.isfinite(d_pos.view(torch.float32, B * C)).all()


def test_kd_mix_guards(pkg, cuda):
    lib = pkg._native.lib()
    n, C = 37, 53
    widths = [24, 12, 24, 24, 12, 24]
    g = torch.Generator().manual_seed(1)
    s = [torch.randn(n, w, generator=g).to(cuda) for w in widths]
    t = [torch.randn(n, w, generator=g).to(cuda) for w in widths]
    sf, tf = torch.randn(n, C, generator=g).to(cuda), torch.randn(n, C, generator=g).to(cuda)
    lab = torch.stack([torch.randint(0, r, (n,), generator=g) for r in (360, 180, 360)], 1).float().to(cuda)
    arr = lambda ts: (ctypes.c_void_p * 6)(*[x.data_ptr() if hasattr(x, "data_ptr") else x for x in ts])
    w_c = (ctypes.c_int32 * 6)(*widths)
    ceb = (ctypes.c_int32 * 3)(15, 15, 15)
    nb = ctypes.c_size_t(0)
    assert lib.crdpn_kd_mix_workspace_bytes(n, ctypes.byref(nb)) == 0
    ws = Guarded(nb.value, cuda, fill=0)          # the ticket word must start zeroed
    loss = Guarded(4, cuda)
    common = (arr(s), arr(t), w_c, sf.data_ptr(), tf.data_ptr(), C, lab.data_ptr(), 3, n, ceb, 15, 0x7FF, 2.0, 0.75, 0.75, 0.25)
    assert lib.crdpn_kd_mix_forward(*common, loss.ptr, ws.ptr, nb.value, _st(cuda)) == 0
    ds = [Guarded(n * w * 4, cuda) for w in widths]
    dt = [Guarded(n * w * 4, cuda) for w in widths]
    dsf, dtf = Guarded(n * C * 4, cuda), Guarded(n * C * 4, cuda)
    rc = lib.crdpn_kd_mix_backward(*common, None, arr([x.ptr for x in ds]), arr([x.ptr for x in dt]), dsf.ptr, dtf.ptr, ws.ptr,
                                   nb.value, _st(cuda))
    assert rc == 0
    torch.cuda.synchronize()
    assert all(x.intact() for x in [ws, loss, dsf, dtf] + ds + dt)
    assert (ws.view(torch.int32, 1) == 0).all()   # the ticket is left zeroed
    for x, w in zip(ds, widths):
        assert torch.isfinite(x.view(torch.float32, n * w)).all()


def test_pointcloud_guards(pkg, cuda):
    lib = pkg._native.lib()
    rng = np.random.default_rng(0)
    counts = [3001, 2500, 7777]
    verts = torch.from_numpy(np.concatenate([rng.normal(size=(v, 3)) for v in counts])).to(cuda)
    offs = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int64, device=cuda)
    B, P = 5, 2500
    ids = torch.tensor([0, 1, 2, 1, 0], dtype=torch.int64, device=cuda)
    rot = torch.tensor([0.0, 10.0, 0.0, 359.0, 45.5], device=cuda)
    out, sub = Guarded(B * 3 * P * 4, cuda), Guarded(B * P * 8, cuda)
    rc = lib.crdpn_pointcloud_sample(verts.data_ptr(), offs.data_ptr(), ids.data_ptr(), rot.data_ptr(), None, 5, 0, B, P, out.ptr,
                                     sub.ptr, _st(cuda))
    assert rc == 0
    torch.cuda.synchronize()
    assert out.intact() and sub.intact()
    o = out.view(torch.float32, B * 3 * P)
    assert o.min() == 0 and o.max() == 1
    s = sub.view(torch.int64, B * P).view(B, P)
    assert all(int(s[b].max()) < counts[int(ids[b])] and s[b].unique().numel() == P for b in range(B))


def test_crd_out_backward_and_stream_step_guards(pkg, cuda):
    lib = pkg._native.lib()
    N, K1, B, D = 5003, 777, 11, 128
    g = torch.Generator().manual_seed(2)
    bank = (torch.rand(N, 2, D, generator=g) - 0.5).to(cuda)
    v1 = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(cuda)
    v2 = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(cuda)
    y = torch.randperm(N, generator=g)[:B].to(cuda)
    idx = torch.randint(0, N, (B, K1), generator=g).to(cuda)
    idx[:, 0] = y
    b1, b2 = bank.data_ptr(), bank.data_ptr() + 4 * D
    # unfused-surface backward
    go = torch.randn(B, K1, generator=g).to(cuda)
    o = torch.rand(B, K1, generator=g).to(cuda)
    old = torch.randn(B, D, generator=g).to(cuda)
    n = ctypes.c_size_t(0)
    assert lib.crdpn_crd_out_backward_workspace_bytes(B, K1, D, ctypes.byref(n)) == 0
    ws, g1, g2 = Guarded(n.value, cuda), Guarded(B * D * 4, cuda), Guarded(B * D * 4, cuda)
    rc = lib.crdpn_crd_out_backward(b1, b2, 2 * D, 0, old.data_ptr(), old.data_ptr(), y.data_ptr(), idx.data_ptr(), go.data_ptr(),
                                    go.data_ptr(), o.data_ptr(), o.data_ptr(), B, K1, D, 0, N, 0.07, g1.ptr, g2.ptr, ws.ptr, n.value,
                                    _st(cuda))
    assert rc == 0
    torch.cuda.synchronize()
    assert ws.intact() and g1.intact() and g2.intact() and torch.isfinite(g1.view(torch.float32, B * D)).all()
    # bank-streaming step
    assert lib.crdpn_crd_stream_workspace_bytes(B, K1, D, N, 0, ctypes.byref(n)) == 0
    ws, g1, g2, res = Guarded(n.value, cuda), Guarded(B * D * 4, cuda), Guarded(B * D * 4, cuda), Guarded(64, cuda)
    rc = lib.crdpn_crd_step(b1, b2, 2 * D, 0, v1.data_ptr(), v2.data_ptr(), idx.data_ptr(), y.data_ptr(), B, K1, D, N, 0, 0, N,
                            0.07, 1000.0, 1200.0, 1e-7, 0.5, 0.5, res.ptr, g1.ptr, g2.ptr, ws.ptr, n.value, 0x200, _st(cuda))
    assert rc == 0, lib.crdpn_last_error()
    torch.cuda.synchronize()
    assert all(x.intact() for x in (ws, g1, g2, res))
    assert res.view(torch.float64, 8)[4].item() == B * K1 and torch.isfinite(g2.view(torch.float32, B * D)).all()
