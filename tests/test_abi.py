"""CPU tests of the C-ABI boundary: the library builds, loads and exports every symbol include/crdpn_b200.h
declares; host-only entry points and argument checking work without a GPU (no kernels are launched)."""
import ctypes

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._native.lib()
    names = pkg._native.declared_symbols()
    assert "crdpn_crd_score" in names and "crdpn_alias_build" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n
        assert n in pkg._native._SIGNATURES, f"{n} has no ctypes signature"
    assert lib.crdpn_abi_version() == 1


def test_host_alias_build_matches_oracle(pkg, oracle):
    lib = pkg._native.lib()
    rng = np.random.default_rng(1)
    for probs in (np.array([0.1, 0.2, 0.3, 0.15, 0.25], np.float32), np.ones(90000, np.float32),
                  rng.random(1001).astype(np.float32)):
        n = probs.size
        prob = np.zeros(n, np.float32)
        alias = np.zeros(n, np.int64)
        rc = lib.crdpn_alias_build(probs.ctypes.data, n, prob.ctypes.data, alias.ctypes.data)
        assert rc == 0
        p2, a2 = oracle.alias_build(probs)
        assert np.array_equal(prob, p2) and np.array_equal(alias, a2)


def test_argument_errors_are_codes_not_crashes(pkg):
    lib = pkg._native.lib()
    assert lib.crdpn_alias_build(None, 5, None, None) == 1000
    rc = lib.crdpn_crd_score(None, None, 128, 0, None, None, None, 1, 1, 128, 10, 0, 0, 10, 0.07, 1.0, 1.0, 1e-7,
                             None, None, None, None, None, None, 0, 0, None)
    assert rc == 1000 and b"null" in lib.crdpn_last_error()
    rc = lib.crdpn_crd_momentum_update(None, None, 128, 0, None, None, None, 1, 128, 0, 10, 0.5, 0.5, None)
    assert rc == 1000
    with pytest.raises(RuntimeError, match="failed"):
        pkg._native.check(rc, "crdpn_crd_momentum_update")


def test_cpu_tensors_are_rejected_loudly(pkg):
    import torch
    opt = type("Opt", (), dict(s_dim=8, t_dim=8, feat_dim=32, n_data=64, nce_k=7, nce_t=0.07, nce_m=0.5))()
    crit = pkg.CRDLoss(opt)
    y = torch.arange(2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.randn(2, 8), torch.randn(2, 8), y, torch.randint(0, 64, (2, 8)))
    sd = crit.state_dict()
    assert {"embed_s.linear.weight", "embed_t.linear.bias", "contrast.params", "contrast.memory_v1",
            "contrast.memory_v2"} <= set(sd)
    assert sd["contrast.memory_v1"].shape == (64, 32) and sd["contrast.params"].tolist()[:2] == [7.0, pytest.approx(0.07)]
    # interleaved [N,2,D] allocation survives load_state_dict
    crit.load_state_dict({k: v.clone() for k, v in sd.items()})
    m1, m2 = crit.contrast.memory_v1, crit.contrast.memory_v2
    assert m1.stride(0) == 64 and m2.data_ptr() - m1.data_ptr() == 32 * 4


def test_product_library_carries_the_sm100a_tensor_core_and_tma_paths(pkg):
    """Static check of the shipped .so (cuobjdump -sass, no GPU): the PointNet, pose-tail and bank-streaming kernels issue
    tcgen05.mma (UTCHMMA) with TMEM loads (LDTM) and bulk / tensor-map copies (UBLKCP / UTMALDG); the gather scoring kernel
    reads rows with 128-bit loads.  A build that silently lost those paths (wrong arch, a fallback body) fails here."""
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    lib_path = pkg._native.LIB_PATH if hasattr(pkg._native, "LIB_PATH") else None
    if lib_path is None:
        from pathlib import Path
        lib_path = Path(pkg.__file__).resolve().parent / "libcrdpn_b200.so"
    elf = subprocess.run(["cuobjdump", "-lelf", str(lib_path)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf   # sm_100a only: no multi-arch fat binary
    sass = subprocess.run(["cuobjdump", "-sass", str(lib_path)], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), set())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur.add(m.group(1))
            cur.add(m.group(1) + m.group(2))

    def ops(fragment):
        hit = [v for k, v in per.items() if fragment in k]
        assert hit, f"no kernel named *{fragment}* in the library"
        return hit

    for frag in ("pointnet_fwd_kernel_v2", "pointnet_fwd_train_split_kernel", "pn_bwd_pass2_kernel", "pose_tail_kernel"):
        for o in ops(frag):
            assert "UTCHMMA" in o and "LDTM" in o and "UBLKCP" in o, (frag, sorted(x for x in o if x.startswith("U")))
    for o in ops("crd_tc_stream_kernel"):
        assert "UTCHMMA" in o and "LDTM" in o and "UTMALDG" in o
    assert any(any(x.startswith("LDG.E.128") or x.startswith("LDG.E.ENL2.256") for x in o) for o in ops("crd_score_kernel"))
