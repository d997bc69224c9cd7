#!/usr/bin/env python
"""Development timing of the large-N Edgewise forward (tcgen05) at the ViT-B/16 core shape."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mop_b200
from mop_b200 import functional as MF


def main():
    B, H, N, dk, V, r = (int(x) for x in (sys.argv[1:7] if len(sys.argv) >= 7 else (256, 12, 196, 64, 5, 4)))
    C = 2 * V + 2
    qkv = torch.randn(B, N, 1, 3, H, dk, device="cuda", dtype=torch.bfloat16)
    sc = [(1 + 0.1 * torch.randn(V, H, 1, dk, device="cuda")) for _ in range(3)]
    head = {"row_proj.weight": torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C), "row_proj.bias": torch.zeros(4 * r, device="cuda"),
            "col_proj.weight": torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C), "col_proj.bias": torch.zeros(4 * r, device="cuda")}
    lg = torch.tensor(-2.0, device="cuda")
    fl = B * H * (2 * N * N * dk * V + 2 * N ** 3 * 2 * (V - 1) + 2 * N * N * dk * 2 + 2 * N * N * 4 * r)
    call = lambda: mop_b200.edgewise_attention(qkv, *sc, lg, head, n_views=V, beta_not=0.5, gate_mode="lowrank", gate_rank=r, impl="tcgen05")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    print(json.dumps(dict(op="edgewise_fwd_large", impl=MF.last_impl["edgewise_fwd"], B=B, H=H, N=N, dk=dk, V=V, r=r, fwd_ms=t,
                          fwd_tflops=fl / t / 1e9, frac_of_sustained_peak=fl / t / 1e9 / 1406.9)))


if __name__ == "__main__":
    main()
