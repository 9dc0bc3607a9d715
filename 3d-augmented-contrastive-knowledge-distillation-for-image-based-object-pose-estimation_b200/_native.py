"""ctypes binding of libcrdpn_b200.so (C ABI in include/crdpn_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc on first use, and if that is
impossible a RuntimeError is raised.  Every wrapper raises RuntimeError on a non-zero return code.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB = None

F32, BF16 = 0, 1


class PoseTailBwdLayer(ctypes.Structure):   # crdpn_pose_tail_bwd_layer (include/crdpn_b200.h)
    _fields_ = [("W", c_void_p), ("y", c_void_p), ("xhat", c_void_p), ("gamma", c_void_p), ("istd", c_void_p), ("g_out", c_void_p),
                ("dW", c_void_p), ("db", c_void_p), ("dgamma", c_void_p), ("dbeta", c_void_p), ("O", c_int64), ("I", c_int64),
                ("src", ctypes.c_int32), ("act", ctypes.c_int32)]


class PoseTailLayer(ctypes.Structure):   # crdpn_pose_tail_layer (include/crdpn_b200.h)
    _fields_ = [("weights", c_void_p), ("bias", c_void_p), ("O", c_int64), ("I", c_int64), ("src", ctypes.c_int32),
                ("act", ctypes.c_int32), ("out", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p),
                ("running_var", c_void_p), ("save_mean", c_void_p), ("save_istd", c_void_p), ("xhat", c_void_p)]


# name -> (restype, argtypes); mirrors include/crdpn_b200.h one-to-one
_SIGNATURES = {
    "crdpn_abi_version": (c_int, []),
    "crdpn_last_error": (ctypes.c_char_p, []),
    "crdpn_launch_count": (c_uint64, []),
    "crdpn_timing_enable": (c_int, [c_int]),
    "crdpn_timing_read": (c_int, [c_int, POINTER(c_double), POINTER(c_uint64)]),
    "crdpn_alias_build": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "crdpn_alias_draw": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_uint64, c_uint64, c_void_p, c_void_p]),
    "crdpn_alias_draw_contrast": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_uint64,
                                          c_uint64, c_void_p, c_void_p]),
    "crdpn_crd_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "crdpn_crd_stream_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "crdpn_crd_score": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                c_float, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_crd_step": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                               c_float, c_float, c_float, c_float, c_float, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_crd_step_drawn": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                     c_float, c_float, c_float, c_float, c_float, c_float,
                                     c_uint64, c_uint64, c_int64, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_crd_loss_forward": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_int,
                                       c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                       c_float, c_float, c_float, c_float, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_alias_draw_contrast_local": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_uint64,
                                                c_uint64, c_void_p, c_void_p]),
    "crdpn_crd_loss_forward_sharded": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64,
                                               c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p,
                                               c_void_p, c_void_p, c_int64, c_int,
                                               c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                               c_float, c_float, c_float, c_float, c_float, c_float,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_crd_step_sharded": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_void_p,
                                       c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                       c_float, c_float, c_float, c_float, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_crd_loss_backward": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_int64] + [c_void_p] * 8),
    "crdpn_crd_momentum_update": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_int64, c_int64, c_int64, c_float, c_float, c_void_p]),
    "crdpn_pointnet_packed_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "crdpn_pointnet_pack": (c_int, [c_void_p] * 18 + [c_float, c_int64, c_void_p, c_void_p]),
    "crdpn_pointnet_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "crdpn_pointnet_forward_eval": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_int, c_void_p]),
    "crdpn_embed_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "crdpn_embed_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "crdpn_nce_kd_workspace_bytes": (c_int, [c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_nce_kd_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_int, c_float, c_uint64,
                                     c_uint64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crdpn_nce_kd_backward": (c_int, [c_void_p, c_int64, c_int64, c_float, c_float, c_uint64, c_uint64, c_void_p, c_size_t,
                                      c_void_p, c_void_p, c_void_p]),
    "crdpn_kd_mix_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "crdpn_kd_mix_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                     c_void_p, ctypes.c_int32, ctypes.c_uint32, c_float, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "crdpn_kd_mix_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                      c_void_p, ctypes.c_int32, ctypes.c_uint32, c_float, c_float, c_float, c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crdpn_pointcloud_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p]),
    "crdpn_crd_out_backward_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_crd_out_backward": (c_int, [c_void_p, c_void_p, c_int64, c_int] + [c_void_p] * 8 +
                               [c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "crdpn_p2p_buffer_bytes": (c_int, [c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "crdpn_p2p_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "crdpn_p2p_free": (c_int, [c_void_p]),
    "crdpn_p2p_export": (c_int, [c_void_p, c_void_p]),
    "crdpn_p2p_import": (c_int, [c_void_p, POINTER(c_void_p)]),
    "crdpn_p2p_close": (c_int, [c_void_p]),
    "crdpn_p2p_allgather_anchors": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                                            c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "crdpn_p2p_allreduce_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64,
                                        c_void_p]),
    "crdpn_p2p_allreduce_blocks": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int64, c_int64, c_void_p]),
    "crdpn_pointnet_train_ctx_bytes": (c_int, [c_int64, c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_pointnet_forward_train": (c_int, [c_void_p, c_int64, c_int64, c_int64] + [c_void_p] * 21 +
                                     [c_float, c_float, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "crdpn_pointnet_forward_train_phased": (c_int, [c_void_p, c_int64, c_int64, c_int64] + [c_void_p] * 21 +
                                            [c_float, c_float, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int64,
                                             c_void_p]),
    "crdpn_pointnet_backward_phased": (c_int, [c_void_p, c_int64, c_int64, c_int64] + [c_void_p] * 9 +
                                       [c_void_p, c_void_p, c_size_t] + [c_void_p] * 12 +
                                       [c_void_p, c_size_t, c_int, c_int, c_int64, c_void_p]),
    "crdpn_pointnet_sync_blocks": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_int), POINTER(c_int),
                                           POINTER(c_size_t), POINTER(c_int64), POINTER(c_int)]),
    "crdpn_pointnet_backward_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_pose_tail_image_bytes": (c_int, [c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_pose_tail_pack_weights": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "crdpn_pose_tail_workspace_bytes": (c_int, [POINTER(PoseTailLayer), c_int, c_int64, c_int64, c_int64, POINTER(c_size_t)]),
    "crdpn_pose_tail_forward": (c_int, [POINTER(PoseTailLayer), c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int,
                                        c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "crdpn_pose_tail_backward_workspace_bytes": (c_int, [POINTER(PoseTailBwdLayer), c_int, c_int64, POINTER(c_size_t)]),
    "crdpn_pose_tail_backward": (c_int, [POINTER(PoseTailBwdLayer), c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "crdpn_pointnet_backward": (c_int, [c_void_p, c_int64, c_int64, c_int64] + [c_void_p] * 9 +
                                [c_void_p, c_void_p, c_size_t] + [c_void_p] * 12 + [c_void_p, c_size_t, c_void_p]),
}


def lib_path() -> Path:
    return _PKG / "libcrdpn_b200.so"


def declared_symbols() -> list[str]:
    """Symbols declared in include/crdpn_b200.h (parsed, so the header stays the single source of truth)."""
    import re
    text = (_PKG.parent / "include" / "crdpn_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crdpn_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        try:
            from importlib import util as _u
            spec = _u.spec_from_file_location("_crdpn_build", _PKG / "build.py")
            mod = _u.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as exc:  # no fallback: fail loudly
            raise RuntimeError(f"libcrdpn_b200.so is missing and could not be built ({exc}); "
                               "this package has no CPU or eager fallback") from exc
    handle = ctypes.CDLL(str(path))
    for name in declared_symbols():
        if not hasattr(handle, name):
            raise RuntimeError(f"{path} does not export {name} (stale build? run build.py --force)")
        if name in _SIGNATURES:
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = _SIGNATURES[name]
    if handle.crdpn_abi_version() != 1:
        raise RuntimeError("libcrdpn_b200.so ABI version mismatch")
    _LIB = handle
    return handle


_DEV = None


def dev_lib() -> ctypes.CDLL:
    """libcrdpn_b200_dev.so: development probes (include/crdpn_b200_dev.h), used by tests / profiling scripts only."""
    global _DEV
    if _DEV is None:
        lib()   # builds both libraries when needed
        _DEV = ctypes.CDLL(str(_PKG / "libcrdpn_b200_dev.so"))
        _DEV.crdpn_umma_tf32_probe.restype = c_int
        _DEV.crdpn_umma_tf32_probe.argtypes = [c_void_p] * 7 + [c_int, c_void_p]
    return _DEV


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().crdpn_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed: {msg}")


def launch_count() -> int:
    return int(lib().crdpn_launch_count())


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def on_device(device):
    """``torch.cuda.device(device)`` only when it is not already the current device (the context manager costs
    ~10 us of host time per call, which matters for a 0.5 ms step made of a dozen launches)."""
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if torch.cuda.current_device() == idx:
        return _NULL
    return torch.cuda.device(idx)


_raw_stream = None


def stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on `device` (the raw getter skips building a Stream object: ~4 us)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream
