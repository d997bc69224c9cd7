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


@pytest.mark.parametrize("B,N,H,dk,R,row0,head,batch,layout", [
    (2, 300, 3, 64, 64, 128, 1, 1, "bnhd"),     # contiguous [B,N,H,dk]
    (2, 196, 4, 56, 128, 128, 2, 0, "qkv"),     # a q/k/v slice of a [B,N,3,H,dk] projection, dk = 56, rows past the end
    (1, 70, 1, 16, 64, 64, 0, 0, "bnhd"),       # single head / batch (degenerate strides), mostly out of bounds
])
def test_tma_tile_load_swizzle128(B, N, H, dk, R, row0, head, batch, layout):
    """cp.async.bulk.tensor tile load through the 4-D tensor map lands as the 128-byte-swizzled operand tile, zero filled."""
    from mop_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(N + dk)
    if layout == "qkv":
        full = torch.randn(B, N, 3, H, dk, generator=g).bfloat16().cuda()
        x = full[:, :, 1]
    else:
        x = torch.randn(B, N, H, dk, generator=g).bfloat16().cuda()
    sb, sn, sh, _ = x.stride()
    out = torch.full((R * 128,), 0xEE, dtype=torch.uint8, device="cuda")
    rc = lib.mop_selftest_tma(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, N, H, dk, C.c_int64(sb), C.c_int64(sn), C.c_int64(sh),
                              R, row0, head, batch, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    got = out.view(torch.bfloat16).view(R, 8, 8).float().cpu()          # [row][stored chunk][8]
    want = torch.zeros(R, 64)
    rows = x[batch, row0:row0 + R, head].float().cpu()
    want[:rows.shape[0], :dk] = rows
    want = want.view(R, 8, 8)                                            # [row][logical chunk][8]
    r = torch.arange(R).view(R, 1)
    stored = torch.arange(8).view(1, 8) ^ (r & 7)                        # logical chunk c of row r is stored at chunk c ^ (r & 7)
    swz = torch.empty_like(want)
    swz[r.expand(R, 8), stored] = want
    assert torch.equal(got, swz)
