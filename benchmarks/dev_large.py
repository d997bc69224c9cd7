#!/usr/bin/env python
"""Development timing of the large-N Edgewise kernels (tcgen05) at the ViT-B/16 core shape: fwd and fwd+bwd."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mop_b200
from mop_b200 import functional as MF


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B, H, N, dk, V, r = (int(x) for x in (args[:6] if len(args) >= 6 else (256, 12, 196, 64, 5, 4)))
    bwd = "--bwd" in sys.argv
    C = 2 * V + 2
    qkv = torch.randn(B, N, 1, 3, H, dk, device="cuda", dtype=torch.bfloat16, requires_grad=bwd)
    sc = [(1 + 0.1 * torch.randn(V, H, 1, dk, device="cuda")).requires_grad_(bwd) for _ in range(3)]
    head = {"row_proj.weight": torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C), "row_proj.bias": torch.zeros(4 * r, device="cuda"),
            "col_proj.weight": torch.randn(4 * r, C, 1, device="cuda") / math.sqrt(C), "col_proj.bias": torch.zeros(4 * r, device="cuda")}
    head = {k: v.requires_grad_(bwd) for k, v in head.items()}
    lg = torch.tensor(-2.0, device="cuda", requires_grad=bwd)
    dy = torch.randn(B, N, H, dk, device="cuda", dtype=torch.bfloat16)
    fl = B * H * (2 * N * N * dk * V + 2 * N ** 3 * 2 * (V - 1) + 2 * N * N * dk * 2 + 2 * N * N * 4 * r)
    call = lambda: mop_b200.edgewise_attention(qkv, *sc, lg, head, n_views=V, beta_not=0.5, gate_mode="lowrank", gate_rank=r, impl="tcgen05")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    MF.kernel_timing = True

    def run():
        if bwd:
            call().backward(dy)
        else:
            with torch.no_grad():
                call()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    MF.kernel_events.clear()
    for _ in range(8):
        flush.zero_()
        run()
    torch.cuda.synchronize()
    out = dict(op="edgewise_large", impl=MF.last_impl["edgewise_fwd"], B=B, H=H, N=N, dk=dk, V=V, r=r)
    for name, evs in MF.kernel_events.items():
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        t = ts[len(ts) // 2]
        mult = 2 if name.endswith("bwd") else 1
        out[name + "_ms"] = t
        out[name + "_tflops"] = mult * fl / t / 1e9
        out[name + "_frac_of_sustained_peak"] = mult * fl / t / 1e9 / 1406.9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
