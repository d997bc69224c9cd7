#!/usr/bin/env python
"""Attention-core microbenchmark over the named BASELINE shapes (M1..M5).

For each shape: forward and forward+backward of the attention CORE through the public functional API
(mop_b200.sdpa / edgewise_attention / quartet_attention), CUDA events on the launching stream, 10 warm-ups,
`--iters` timed iterations, L2 flushed between iterations (256 MiB memset outside the event pairs).
Prints one JSON line per (shape, impl): algorithmic TFLOP/s (dense contractions only, bwd = 2 x fwd) and the
fraction of the measured burst bf16 peak (kernels timed in isolation).
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mop_b200  # noqa: E402
from mop_b200 import functional as MF  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]), "measured burst"
    except Exception:
        return 1590.0, "fallback"


def timeit(fn, iters, flush, graph=True):
    """Median device time of fn().  The launches of one call are captured in a CUDA graph and the replays are timed (the
    training step replays graphs too): for the small shapes an eager call is bound by the ~10 host-side launches, not by
    the kernels.  Falls back to eager launches if the capture fails."""
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    run = fn
    if graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run = g.replay
            for _ in range(3):
                run()
            torch.cuda.synchronize()
        except Exception:
            torch.cuda.synchronize()
            run = fn
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    timeit.last_launch = "cuda graph replay" if run is not fn else "eager"
    return ts[len(ts) // 2]


def sdpa_case(name, B, H, N, dk, causal, impl, iters, flush, dtype=torch.bfloat16):
    q, k, v = (torch.randn(B, N, H, dk, device="cuda", dtype=dtype, requires_grad=True) for _ in range(3))
    dy = torch.randn(B, N, H, dk, device="cuda", dtype=dtype)
    fl = B * H * 4 * N * N * dk / (2 if causal else 1)
    fwd = lambda: mop_b200.sdpa(q, k, v, causal=causal, impl=impl)

    def fb():
        y = mop_b200.sdpa(q, k, v, causal=causal, impl=impl)
        y.backward(dy)
        q.grad = k.grad = v.grad = None
    with torch.no_grad():
        t_f = timeit(fwd, iters, flush)
    t_fb = timeit(fb, iters, flush)
    return dict(shape=name, op="sdpa", impl=MF.last_impl["sdpa_fwd"], B=B, H=H, N=N, dk=dk, causal=causal, dtype=str(dtype),
                fwd_ms=t_f, fwd_bwd_ms=t_fb, fwd_tflops=fl / t_f / 1e9, fwd_bwd_tflops=3 * fl / t_fb / 1e9,
                launch=timeit.last_launch)


def edgewise_case(name, B, H, N, dk, V, r, impl, iters, flush, dtype=torch.bfloat16):
    C = 2 * V + 2
    qkv = torch.randn(B, N, 1, 3, H, dk, device="cuda", dtype=dtype, requires_grad=True)
    sc = [(1 + 0.1 * torch.randn(V, H, 1, dk, device="cuda")).requires_grad_(True) for _ in range(3)]
    head = {"row_proj.weight": (torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C)).requires_grad_(True),
            "row_proj.bias": torch.zeros(4 * r, device="cuda", requires_grad=True),
            "col_proj.weight": (torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C)).requires_grad_(True),
            "col_proj.bias": torch.zeros(4 * r, device="cuda", requires_grad=True)}
    lg = torch.tensor(-2.0, device="cuda", requires_grad=True)
    dy = torch.randn(B, N, H, dk, device="cuda", dtype=dtype)
    fl = B * H * (2 * N * N * dk * V + 2 * N ** 3 * 2 * (V - 1) + 2 * N * N * dk * 2 + 2 * N * N * 4 * r)
    call = lambda: mop_b200.edgewise_attention(qkv, *sc, lg, head, n_views=V, beta_not=0.5, gate_mode="lowrank", gate_rank=r, impl=impl)

    def fb():
        call().backward(dy)
    with torch.no_grad():
        t_f = timeit(call, iters, flush)
    t_fb = timeit(fb, iters, flush)
    return dict(shape=name, op="edgewise", impl=MF.last_impl["edgewise_fwd"], B=B, H=H, N=N, dk=dk, V=V, r=r, dtype=str(dtype),
                fwd_ms=t_f, fwd_bwd_ms=t_fb, fwd_tflops=fl / t_f / 1e9, fwd_bwd_tflops=3 * fl / t_fb / 1e9,
                launch=timeit.last_launch)


def quartet_case(name, B, H, T, dk, impl, iters, flush, dtype=torch.bfloat16):
    ts = [torch.randn(B, T, H, dk, device="cuda", dtype=dtype, requires_grad=True) for _ in range(5)]
    mix = torch.tensor([-5.0], device="cuda", requires_grad=True)
    gam = torch.tensor([1.0], device="cuda", requires_grad=True)
    dy = torch.randn(B, T, H, dk, device="cuda", dtype=dtype)
    fl = B * H * 3 * T * T * dk
    call = lambda: mop_b200.quartet_attention(ts[0], ts[1], ts[2], ts[3], ts[4], mix, gam, impl=impl)

    def fb():
        call().backward(dy)
    with torch.no_grad():
        t_f = timeit(call, iters, flush)
    t_fb = timeit(fb, iters, flush)
    return dict(shape=name, op="quartet", impl=MF.last_impl["quartet_fwd"], B=B, H=H, N=T, dk=dk, dtype=str(dtype),
                fwd_ms=t_f, fwd_bwd_ms=t_fb, fwd_tflops=fl / t_f / 1e9, fwd_bwd_tflops=3 * fl / t_fb / 1e9,
                launch=timeit.last_launch)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pk, src = peak()
    rows = []
    rows.append(edgewise_case("M1/M2 Edgewise E (config 1/2 core)", 256, 4, 64, 56, 5, 4, None, a.iters, flush))
    rows.append(edgewise_case("M3 Edgewise ViT-B/16 core", 256, 12, 196, 64, 5, 4, "tcgen05", max(3, a.iters // 2), flush))
    rows.append(quartet_case("M4 Quartet GPT-1024 (B=16)", 16, 12, 1024, 64, "tcgen05", a.iters, flush))
    rows.append(quartet_case("M4 Quartet GPT-4096 (B=4)", 4, 12, 4096, 64, "tcgen05", a.iters, flush))
    rows.append(sdpa_case("M1 MSA model A (config 1)", 256, 4, 64, 56, False, "tcgen05", a.iters, flush))
    rows.append(sdpa_case("M3 ViT-B/16 plain attention", 256, 12, 196, 64, False, "tcgen05", a.iters, flush))
    rows.append(sdpa_case("M5 Whisper encoder self-attention", 8, 16, 1500, 64, False, "tcgen05", a.iters, flush))
    rows.append(sdpa_case("GPT-1024 causal plain attention", 16, 12, 1024, 64, True, "tcgen05", a.iters, flush))
    if not a.quick:
        # the fp32-mode (CUDA-core) kernels on the same bf16 storage, reduced batch: what the tcgen05 paths replace
        rows.append(edgewise_case("M3 Edgewise ViT-B/16 core (fp32-mode kernels, B=8)", 8, 12, 196, 64, 5, 4, "simt", 3, flush))
        rows.append(quartet_case("M4 Quartet GPT-1024 (fp32-mode kernels, B=2)", 2, 12, 1024, 64, "simt", 3, flush))
    for r in rows:
        r["frac_of_bf16_peak_fwd_bwd"] = r["fwd_bwd_tflops"] / pk
        r["peak"] = f"{pk} TFLOP/s ({src})"
        r["l2"] = "flushed between iterations"
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
