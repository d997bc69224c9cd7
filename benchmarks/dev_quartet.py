#!/usr/bin/env python
"""Development timing of the Quartet attention core (fwd, bwd) through the functional API."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mop_b200
from mop_b200 import functional as MF

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, H, T, dk = (int(x) for x in (args[:4] if len(args) >= 4 else (16, 12, 1024, 64)))
impl = "simt" if "--simt" in sys.argv else "tcgen05"
ts = [torch.randn(B, T, H, dk, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(5)]
mix = torch.tensor([-5.0], device="cuda", requires_grad=True)
gam = torch.tensor([1.0], device="cuda", requires_grad=True)
dy = torch.randn(B, T, H, dk, device="cuda", dtype=torch.bfloat16)
fl = B * H * 3 * T * T * dk
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
MF.kernel_timing = True
run = lambda: mop_b200.quartet_attention(*ts, mix, gam, impl=impl).backward(dy)
for _ in range(3):
    run()
torch.cuda.synchronize()
MF.kernel_events.clear()
for _ in range(8):
    flush.zero_(); run()
torch.cuda.synchronize()
out = dict(op="quartet", impl=MF.last_impl["quartet_fwd"], B=B, H=H, T=T, dk=dk)
for name, evs in MF.kernel_events.items():
    t = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    mult = 2 if name.endswith("bwd") else 1
    out[name + "_ms"] = t
    out[name + "_tflops"] = mult * fl / t / 1e9
    out[name + "_frac_of_sustained_peak"] = mult * fl / t / 1e9 / 1406.9
print(json.dumps(out))
