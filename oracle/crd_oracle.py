"""oracle/crd_oracle.py -- CPU oracle for the CRD memory-bank NCE step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does.

PARITY UNPINNED: ``/root/reference`` ships no CRD memory-bank code and
no tests or golden vectors for it; the algorithm lives in the un-vendored, un-pinned third-party module
HobbitLong/RepDistiller (``crd/criterion.py``, ``crd/memory.py``).  Two restatements of that *published*
algorithm live here:

* ``libcrd_oracle.so`` (``crd_oracle.c``): fp64-accumulating scorer, canonical-order momentum update,
  Vose alias tables and the Philox draw -- the checker the CUDA kernels are compared with.
* ``StockCRD``: the stock ``index_select -> bmm -> exp -> /Z -> NCE loss -> autograd -> index_copy_``
  formulation in fp32 torch ops on CPU -- the formulation a user of the public CRD code would run, used
  as the timed CPU baseline ("port") and as an independent autograd cross-check of the closed-form
  gradients in the C file.

Insertion point in the reference: next to ``representation_loss`` in ``calculate_kd_loss_new``
(``KD/vision/vanilla/vanilla_kd.py:158-160``), called from ``KD/common/base_class.py:387``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    """Compile crd_oracle.c with gcc (see Makefile). Returns the .so path."""
    so = _HERE / "_build" / "libcrd_oracle.so"
    src = _HERE / "crd_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B"] if force else ["make", "-C", str(_HERE), "-s"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.oracle_alias_build.restype = ctypes.c_int
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def philox(seed: int, ctr: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox(ctypes.c_uint64(seed), ctypes.c_uint64(ctr), _p(out))
    return out


def alias_build(probs: np.ndarray):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    n = probs.shape[0]
    prob = np.zeros(n, dtype=np.float32)
    alias = np.zeros(n, dtype=np.int64)
    rc = lib().oracle_alias_build(_p(probs), ctypes.c_int64(n), _p(prob), _p(alias))
    if rc != 0:
        raise RuntimeError(f"oracle_alias_build rc={rc}")
    return prob, alias


def alias_draw(prob: np.ndarray, alias: np.ndarray, count: int, seed: int, offset: int = 0) -> np.ndarray:
    out = np.zeros(count, dtype=np.int64)
    lib().oracle_alias_draw(_p(prob), _p(alias), ctypes.c_int64(prob.shape[0]), ctypes.c_int64(count),
                            ctypes.c_uint64(seed), ctypes.c_uint64(offset), _p(out))
    return out


def alias_draw_contrast(prob, alias, y: np.ndarray, K1: int, seed: int, offset: int = 0) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.int64)
    B = y.shape[0]
    out = np.zeros((B, K1), dtype=np.int64)
    lib().oracle_alias_draw_contrast(_p(prob), _p(alias), ctypes.c_int64(prob.shape[0]), _p(y),
                                     ctypes.c_int64(B), ctypes.c_int64(K1), ctypes.c_uint64(seed),
                                     ctypes.c_uint64(offset), _p(out))
    return out


def crd_score(bank1: np.ndarray, bank2: np.ndarray, v1: np.ndarray, v2: np.ndarray, idx: np.ndarray,
              n_data: int, T: float, Z1: float, Z2: float, eps: float = 1e-7,
              row_begin: int = 0, row_end: int | None = None, want_out: bool = True, k_total: int = 0):
    """fp64 scorer. bank1/bank2: [N_local, D] fp32 (any row stride, last dim contiguous).
    Returns dict(loss_s, loss_t, sum_e1, sum_e2, count, out_v1, out_v2, grad_v1, grad_v2)."""
    assert bank1.dtype == np.float32 and bank2.dtype == np.float32
    assert bank1.strides[1] == 4 and bank2.strides[1] == 4 and bank1.strides[0] == bank2.strides[0]
    v1 = np.ascontiguousarray(v1, dtype=np.float32)
    v2 = np.ascontiguousarray(v2, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, K1 = idx.shape
    D = v1.shape[1]
    if row_end is None:
        row_end = row_begin + bank1.shape[0]
    out1 = np.zeros((B, K1), dtype=np.float64) if want_out else None
    out2 = np.zeros((B, K1), dtype=np.float64) if want_out else None
    res = np.zeros(5, dtype=np.float64)
    g1 = np.zeros((B, D), dtype=np.float64)
    g2 = np.zeros((B, D), dtype=np.float64)
    lib().oracle_crd_score(_p(bank1), _p(bank2), ctypes.c_int64(bank1.strides[0] // 4), _p(v1), _p(v2), _p(idx),
                           ctypes.c_int64(B), ctypes.c_int64(K1), ctypes.c_int64(D), ctypes.c_int64(n_data),
                           ctypes.c_int64(k_total), ctypes.c_int64(row_begin), ctypes.c_int64(row_end),
                           ctypes.c_double(T), ctypes.c_double(Z1), ctypes.c_double(Z2), ctypes.c_double(eps),
                           _p(out1) if want_out else None, _p(out2) if want_out else None,
                           _p(res), _p(g1), _p(g2))
    return dict(loss_s=res[0], loss_t=res[1], sum_e1=res[2], sum_e2=res[3], count=res[4],
                out_v1=out1, out_v2=out2, grad_v1=g1, grad_v2=g2)


def momentum_update(bank: np.ndarray, v: np.ndarray, y: np.ndarray, m: float,
                    row_begin: int = 0, row_end: int | None = None) -> None:
    """In-place canonical-order momentum update of one bank (fp32, last dim contiguous)."""
    assert bank.dtype == np.float32 and bank.strides[1] == 4
    v = np.ascontiguousarray(v, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.int64)
    B, D = v.shape
    if row_end is None:
        row_end = row_begin + bank.shape[0]
    m32 = np.float32(m)
    om = np.float32(1.0 - float(m32))  # python-float (1 - momentum) cast to fp32 by the scalar multiply
    lib().oracle_momentum_update(_p(bank), ctypes.c_int64(bank.strides[0] // 4), _p(v), _p(y),
                                 ctypes.c_int64(B), ctypes.c_int64(D), ctypes.c_int64(row_begin),
                                 ctypes.c_int64(row_end), ctypes.c_float(m32), ctypes.c_float(om))


def l2_normalize(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().oracle_l2_normalize(_p(x), _p(out), ctypes.c_int64(x.shape[0]), ctypes.c_int64(x.shape[1]))
    return out


def bank_init(n_data: int, dim: int, seed: int) -> np.ndarray:
    """Bank init of the published algorithm: U(-s, s), s = 1/sqrt(dim/3)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    stdv = 1.0 / math.sqrt(dim / 3.0)
    return torch.rand(n_data, dim, generator=g).mul_(2 * stdv).add_(-stdv).numpy()


# ---------------------------------------------------------------------------------------------------------
# Stock formulation in torch ops (CPU fp32): the timed baseline and the autograd cross-check.
# ---------------------------------------------------------------------------------------------------------
class StockCRD:
    """index_select -> bmm -> exp -> /Z -> NCE -> backward -> index_copy_  (published CRD, stock form).

    Holds two banks, the frozen normalisers Z and the embed heads' parameters.  Pure torch ops so that
    autograd supplies the gradients independently of the closed form used by the C oracle / CUDA kernel.
    """

    def __init__(self, s_dim, t_dim, feat_dim, n_data, nce_k, nce_t=0.07, nce_m=0.5, seed=46, dtype=None):
        import torch
        self.torch = torch
        dtype = dtype or torch.float32
        g = torch.Generator().manual_seed(seed)
        bs, bt = 1.0 / math.sqrt(s_dim), 1.0 / math.sqrt(t_dim)
        self.Ws = ((torch.rand(feat_dim, s_dim, generator=g) * 2 - 1) * bs).to(dtype).requires_grad_()
        self.bs = ((torch.rand(feat_dim, generator=g) * 2 - 1) * bs).to(dtype).requires_grad_()
        self.Wt = ((torch.rand(feat_dim, t_dim, generator=g) * 2 - 1) * bt).to(dtype).requires_grad_()
        self.bt = ((torch.rand(feat_dim, generator=g) * 2 - 1) * bt).to(dtype).requires_grad_()
        stdv = 1.0 / math.sqrt(feat_dim / 3.0)
        self.memory_v1 = (torch.rand(n_data, feat_dim, generator=g) * 2 * stdv - stdv).to(dtype)
        self.memory_v2 = (torch.rand(n_data, feat_dim, generator=g) * 2 * stdv - stdv).to(dtype)
        self.K, self.T, self.m, self.n_data = nce_k, nce_t, nce_m, n_data
        self.Z1 = -1.0
        self.Z2 = -1.0
        self.eps = 1e-7

    def embed(self, x, W, b):
        x = x.reshape(x.shape[0], -1)
        x = self.torch.nn.functional.linear(x, W, b)
        return x / x.pow(2).sum(1, keepdim=True).pow(0.5)

    def contrast(self, v1, v2, y, idx):
        torch = self.torch
        B, K1, D = v1.shape[0], self.K + 1, self.memory_v1.shape[1]
        w1 = torch.index_select(self.memory_v1, 0, idx.reshape(-1)).detach().view(B, K1, D)
        out_v2 = torch.exp(torch.bmm(w1, v2.view(B, D, 1)) / self.T)
        w2 = torch.index_select(self.memory_v2, 0, idx.reshape(-1)).detach().view(B, K1, D)
        out_v1 = torch.exp(torch.bmm(w2, v1.view(B, D, 1)) / self.T)
        if self.Z1 < 0:
            self.Z1 = float(out_v1.mean().item() * self.n_data)
        if self.Z2 < 0:
            self.Z2 = float(out_v2.mean().item() * self.n_data)
        out_v1 = out_v1 / self.Z1
        out_v2 = out_v2 / self.Z2
        with torch.no_grad():
            for mem, v in ((self.memory_v1, v1), (self.memory_v2, v2)):
                pos = torch.index_select(mem, 0, y.view(-1))
                pos.mul_(self.m)
                pos.add_(v * (1 - self.m))
                nrm = pos.pow(2).sum(1, keepdim=True).pow(0.5)
                mem.index_copy_(0, y, pos / nrm)
        return out_v1, out_v2

    def nce(self, x):
        torch = self.torch
        B, m = x.shape[0], x.shape[1] - 1
        Pn = 1.0 / float(self.n_data)
        pos = x.select(1, 0)
        log_d1 = (pos / (pos + (m * Pn + self.eps))).log()
        neg = x.narrow(1, 1, m)
        log_d0 = (torch.full_like(neg, m * Pn) / (neg + (m * Pn + self.eps))).log()
        return -(log_d1.sum(0) + log_d0.reshape(-1, 1).sum(0)) / B

    def loss(self, f_s, f_t, y, contrast_idx):
        v1 = self.embed(f_s, self.Ws, self.bs)
        v2 = self.embed(f_t, self.Wt, self.bt)
        o1, o2 = self.contrast(v1, v2, y, contrast_idx)
        return (self.nce(o1) + self.nce(o2)).reshape(())

    def step(self, f_s, f_t, y, contrast_idx):
        """One full forward + backward + bank update; returns the loss value."""
        for p in (self.Ws, self.bs, self.Wt, self.bt):
            p.grad = None
        l = self.loss(f_s, f_t, y, contrast_idx)
        l.backward()
        return l.detach()
