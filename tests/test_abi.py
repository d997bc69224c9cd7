"""CPU checks of the C-ABI library: it builds, loads, exports every symbol the header declares,
agrees with the ctypes structs on their size, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import mop_b200.build as b
    b.build()
    from mop_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "mop_b200.h")).read()
    names = set(re.findall(r"\b(mop_[a-z0-9_]+)\s*\(", hdr))
    assert {"mop_abi_version", "mop_last_error", "mop_edgewise_fwd", "mop_edgewise_bwd", "mop_sdpa_fwd",
            "mop_sdpa_bwd", "mop_quartet_fwd", "mop_quartet_bwd"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_struct_sizes_agree(lib):
    from mop_b200 import _lib
    p = _lib.new_params(_lib.EdgewiseParams)
    p.B, p.H, p.N, p.dk, p.V, p.Vp, p.gate_mode, p.gate_rank, p.hidden = 2, 2, 64, 56, 5, 1, 1, 4, 16
    p.qkv = p.y = p.chain_value_logit = p.row_w = p.row_b = p.col_w = p.col_b = 1
    assert lib.mop_edgewise_workspace_bytes(C.byref(p), 0) > 0, _lib.last_error()
    assert lib.mop_edgewise_workspace_bytes(C.byref(p), 1) > lib.mop_edgewise_workspace_bytes(C.byref(p), 0)
    p.struct_bytes += 8
    assert lib.mop_edgewise_workspace_bytes(C.byref(p), 0) == 0
    assert "size mismatch" in _lib.last_error()
    s = _lib.new_params(_lib.SdpaParams)
    s.B, s.H, s.Nq, s.Nk, s.dk = 1, 1, 64, 64, 64
    s.q = s.k = s.v = s.y = 1
    assert lib.mop_sdpa_workspace_bytes(C.byref(s), 1) > 0, _lib.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback(lib):
    import mop_b200
    m = mop_b200.EdgewiseMSA(16, heads=2, n_views=2, share_qkv=True, gate_mode="lowrank")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        mop_b200.MSA(16, heads=2)(torch.randn(1, 4, 16))
    from mop_b200 import _lib
    p = _lib.new_params(_lib.EdgewiseParams)
    p.B, p.H, p.N, p.dk, p.V, p.Vp, p.gate_mode, p.gate_rank, p.hidden = 1, 1, 4, 8, 2, 1, 1, 4, 16
    p.qkv = p.y = p.chain_value_logit = p.row_w = p.row_b = p.col_w = p.col_b = 1
    assert lib.mop_edgewise_fwd(C.byref(p), None) == -4  # MOP_ECUDA
