"""ctypes binding of libmop_b200.so (the C ABI declared in include/mop_b200.h).

The library is built in-tree by ``mop_b200.build.build()`` (nvcc, sm_100a).
There is no CPU fallback: if the shared object is missing or a call fails the
caller gets a ``RuntimeError`` carrying ``mop_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOP_B200_LIB") or os.path.join(_HERE, "libmop_b200.so")   # override: experiment builds only

MOP_ABI_VERSION = 9
MOP_F32, MOP_BF16 = 0, 1
MOP_GATE_DENSE, MOP_GATE_LOWRANK, MOP_GATE_CONST = 0, 1, 2
MOP_IMPL_AUTO, MOP_IMPL_SIMT, MOP_IMPL_TCGEN05 = 0, 1, 2
IMPL_NAMES = {MOP_IMPL_SIMT: "simt", MOP_IMPL_TCGEN05: "tcgen05"}

i32, i64, f32, vp, sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class EdgewiseParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("dtype", i32), ("impl", i32), ("impl_used", i32),
        ("B", i32), ("H", i32), ("N", i32), ("dk", i32), ("V", i32), ("Vp", i32),
        ("gate_mode", i32), ("gate_rank", i32), ("hidden", i32), ("use_k3", i32),
        ("beta_not", f32), ("eps", f32),
        ("qkv", vp), ("y", vp),
        ("q_scale", vp), ("k_scale", vp), ("v_scale", vp), ("chain_value_logit", vp),
        ("row_w", vp), ("row_b", vp), ("col_w", vp), ("col_b", vp),
        ("conv1_w", vp), ("conv1_b", vp), ("mid3_w", vp), ("mid3_b", vp), ("conv2_w", vp), ("conv2_b", vp),
        ("row_stats", vp), ("y_base", vp), ("aux", vp),
        ("dy", vp), ("dqkv", vp), ("dscale_part", vp), ("dhead_part", vp), ("dlogit_part", vp),
        ("workspace", vp), ("workspace_bytes", sz),
        ("const_gates", f32 * 4), ("hops", i32),
        ("lens_n", i32), ("lens_dil", i32 * 4), ("lens_w", vp), ("dlens_part", vp),
    ]


class SdpaParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("dtype", i32), ("impl", i32), ("impl_used", i32),
        ("B", i32), ("H", i32), ("Nq", i32), ("Nk", i32), ("dk", i32), ("causal", i32), ("scale", f32),
        ("q", vp), ("k", vp), ("v", vp),
        ("q_sb", i64), ("q_sn", i64), ("q_sh", i64), ("k_sb", i64), ("k_sn", i64), ("k_sh", i64),
        ("v_sb", i64), ("v_sn", i64), ("v_sh", i64),
        ("y", vp),
        ("bias", vp), ("bias_sb", i64), ("bias_sh", i64), ("bias_sq", i64), ("bias_sk", i64),
        ("zero_mask", vp), ("zm_sb", i64), ("zm_sh", i64), ("zm_sq", i64), ("zm_sk", i64),
        ("lse", vp),
        ("dy", vp), ("dq", vp), ("dk_", vp), ("dv", vp),
        ("workspace", vp), ("workspace_bytes", sz),
        ("dropout_p", f32), ("dropout_seed", C.c_uint64), ("dropout_offset", C.c_uint64),
    ]


class QuartetParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("dtype", i32), ("impl", i32), ("impl_used", i32),
        ("B", i32), ("H", i32), ("T", i32), ("dk", i32), ("use_quartet", i32), ("scale", f32), ("eps", f32),
        ("q", vp), ("k", vp), ("v", vp), ("q2", vp), ("k2", vp),
        ("mixture", vp), ("quartet_scale", vp),
        ("add_mask", vp), ("am_sb", i64), ("am_sh", i64), ("am_sq", i64), ("am_sk", i64),
        ("y", vp), ("stats", vp), ("y_f32", vp),
        ("dy", vp), ("dq", vp), ("dk_", vp), ("dv", vp), ("dq2", vp), ("dk2", vp),
        ("dscalar_part", vp),
        ("workspace", vp), ("workspace_bytes", sz),
        ("dropout_p", f32), ("dropout_seed", C.c_uint64), ("dropout_offset", C.c_uint64),
        ("fwd_workspace", vp), ("fwd_workspace_bytes", sz),
    ]


class TokenGateParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("dtype", i32), ("B", i32), ("T", i32), ("D", i32), ("Gh", i32), ("Gw", i32),
        ("V", i32), ("K", i32), ("hid", i32), ("nparts", i32),
        ("x", vp), ("views_w", vp), ("k3_w", vp), ("k1_w", vp), ("f1_w", vp), ("f2_w", vp), ("f2_b", vp), ("a_pos", vp), ("a_neg", vp),
        ("out", vp), ("views", vp), ("gate", vp), ("dout", vp), ("dx", vp), ("dwv_part", vp), ("dnet_part", vp),
    ]


class TokenGate1dParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("dtype", i32), ("B", i32), ("T", i32), ("D", i32), ("V", i32), ("nparts", i32),
        ("x", vp), ("views_w", vp), ("w_eff", vp), ("out", vp), ("views", vp), ("gate", vp), ("dout", vp), ("dx", vp),
        ("dwv_part", vp), ("dweff_part", vp),
    ]


class LnParams(C.Structure):
    _fields_ = [
        ("struct_bytes", i32), ("rows", i32), ("D", i32), ("r_dtype", i32), ("y_dtype", i32), ("rows_per_sample", i32),
        ("nparts", i32), ("eps", f32),
        ("x", vp), ("r", vp), ("scale", vp), ("gamma", vp), ("beta", vp),
        ("x_new", vp), ("y", vp), ("mean", vp), ("rstd", vp),
        ("dy", vp), ("dx_new", vp), ("dx", vp), ("dr", vp), ("dgamma_part", vp), ("dbeta_part", vp),
    ]


_lock = threading.Lock()
_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import mop_b200.build as b; b.build()'` "
                "(nvcc, sm_100a).  mop_b200 has no CPU/PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.mop_abi_version.restype = C.c_int
        lib.mop_last_error.restype = C.c_char_p
        lib.mop_device_sm_count.restype = C.c_int
        for name, st in (("edgewise", EdgewiseParams), ("sdpa", SdpaParams), ("quartet", QuartetParams)):
            ws = getattr(lib, f"mop_{name}_workspace_bytes")
            ws.restype = C.c_size_t
            ws.argtypes = [C.POINTER(st), C.c_int]
            for d in ("fwd", "bwd"):
                fn = getattr(lib, f"mop_{name}_{d}")
                fn.restype = C.c_int
                fn.argtypes = [C.POINTER(st), C.c_void_p]
        lib.mop_ln_partial_rows_d.restype = C.c_int
        lib.mop_ln_partial_rows_d.argtypes = [C.c_int, C.c_int]
        lib.mop_edgewise_reduce_partials.restype = C.c_int
        lib.mop_edgewise_reduce_partials.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mop_token_gate_partial_rows.restype = C.c_int
        lib.mop_token_gate_partial_rows.argtypes = [C.c_int]
        lib.mop_token_gate_wv_groups.restype = C.c_int
        lib.mop_token_gate_wv_groups.argtypes = [C.c_int]
        lib.mop_token_gate_net_params.restype = C.c_int
        lib.mop_token_gate_net_params.argtypes = [C.c_void_p]
        for fn in (lib.mop_token_gate_fwd, lib.mop_token_gate_bwd):
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p]
        lib.mop_token_gate1d_partial_rows.restype = C.c_int
        lib.mop_token_gate1d_partial_rows.argtypes = [C.c_int, C.c_int]
        for fn in (lib.mop_token_gate1d_fwd, lib.mop_token_gate1d_bwd):
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p]
        lib.mop_mop2d_partial_rows.restype = C.c_int
        lib.mop_mop2d_fwd.restype = C.c_int
        lib.mop_mop2d_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.mop_mop2d_bwd.restype = C.c_int
        lib.mop_mop2d_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.mop_dropout_mask.restype = C.c_int
        lib.mop_dropout_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.mop_ln_partial_rows.restype = C.c_int
        lib.mop_ln_partial_rows.argtypes = [C.c_int]
        for d in ("fwd", "bwd"):
            fn = getattr(lib, f"mop_ln_{d}")
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(LnParams), C.c_void_p]
        lib.mop_edgewise_needs_row_stats.restype = C.c_int
        lib.mop_edgewise_needs_row_stats.argtypes = [C.POINTER(EdgewiseParams)]
        lib.mop_edgewise_partial_rows.restype = C.c_int
        lib.mop_edgewise_partial_rows.argtypes = [C.POINTER(EdgewiseParams)]
        lib.mop_edgewise_aux_floats.restype = C.c_size_t
        lib.mop_edgewise_aux_floats.argtypes = [C.POINTER(EdgewiseParams)]
        lib.mop_edgewise_head_param_count.restype = C.c_size_t
        lib.mop_edgewise_head_param_count.argtypes = [C.POINTER(EdgewiseParams)]
        if lib.mop_abi_version() != MOP_ABI_VERSION:
            raise RuntimeError(f"libmop_b200 ABI {lib.mop_abi_version()} != binding {MOP_ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def last_error() -> str:
    return load().mop_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def new_params(cls):
    p = cls()
    p.struct_bytes = C.sizeof(cls)
    return p
