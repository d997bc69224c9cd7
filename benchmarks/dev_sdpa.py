#!/usr/bin/env python
"""Development timing of the plain attention core (fwd, bwd) through the functional API: dev_sdpa.py B H N dk [--causal]."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mop_b200
from mop_b200 import functional as MF

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, H, N, dk = (int(x) for x in (args[:4] if len(args) >= 4 else (8, 16, 1500, 64)))
causal = "--causal" in sys.argv
q, k, v = (torch.randn(B, N, H, dk, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
dy = torch.randn(B, N, H, dk, device="cuda", dtype=torch.bfloat16)
fl = B * H * 4 * N * N * dk // (2 if causal else 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
MF.kernel_timing = True
run = lambda: mop_b200.sdpa(q, k, v, causal=causal, impl="tcgen05").backward(dy)
for _ in range(3):
    run()
torch.cuda.synchronize()
MF.kernel_events.clear()
for _ in range(8):
    flush.zero_(); run()
torch.cuda.synchronize()
out = dict(op="sdpa", B=B, H=H, N=N, dk=dk, causal=causal)
for name, evs in MF.kernel_events.items():
    t = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    mult = 2 if name.endswith("bwd") else 1
    out[name + "_ms"] = t
    out[name + "_tflops"] = mult * fl / t / 1e9
print(json.dumps(out))
