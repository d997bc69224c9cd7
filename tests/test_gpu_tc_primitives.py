"""tcgen05 / TMEM building blocks (descriptors, M=64 MMA, lane-interleaved accumulators, 16x256b ld/st)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("lane_off,col_off", [(0, 0), (16, 0), (0, 64), (16, 32)])
def test_umma_64x64x64(a_mn, b_mn, lane_off, col_off):
    from mop_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(a_mn * 2 + b_mn)
    A = torch.randn(64, 64, generator=g).bfloat16().float()
    B = torch.randn(64, 64, generator=g).bfloat16().float()
    ref = (A.double().T if a_mn else A.double()) @ (B.double() if b_mn else B.double().T)
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((64, 64), float("nan"), device="cuda")
    D2 = torch.full((64, 64), float("nan"), device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.mop_selftest_umma(p(Ad), p(Bd), p(D), p(D2), a_mn, b_mn, lane_off, col_off,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert (D.double().cpu() - ref).abs().max().item() < 1e-4
    assert torch.equal(D2, D + 1.0)


# (Ma, Nn, K, b_mn, Ra, Rb, Kb, b_k0): the operand shapes of edgewise_tc_large.cuh
M128_CASES = [
    (196, 208, 64, 0, 200, 208, 0, 0),      # S_k = Q Ks^T
    (196, 208, 208, 1, 200, 208, 208, 0),   # chain product X A_k
    (200, 32, 64, 0, 200, 32, 0, 0),        # score panel
    (128, 64, 32, 1, 128, 208, 208, 160),   # P V_1 with a row offset into the value tile
    (100, 112, 112, 1, 200, 208, 112, 0),
]


@pytest.mark.parametrize("Ma,Nn,K,b_mn,Ra,Rb,Kb,b_k0", M128_CASES)
def test_umma_m128_thread_per_row(Ma, Nn, K, b_mn, Ra, Rb, Kb, b_k0):
    from mop_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(Ma + Nn + K)
    A = torch.randn(Ma, K, generator=g).bfloat16().float()
    B = (torch.randn(Kb, Nn, generator=g) if b_mn else torch.randn(Nn, K, generator=g)).bfloat16().float()
    ref = A.double() @ (B.double()[b_k0:b_k0 + K] if b_mn else B.double().T)
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((256, Nn), float("nan"), device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.mop_selftest_umma128(p(Ad), p(Bd), p(D), Ma, Nn, K, b_mn, Ra, Rb, Kb, b_k0,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    err = (D[:Ma].double().cpu() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
