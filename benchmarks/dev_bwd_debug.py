#!/usr/bin/env python
"""Development: per-gradient errors of the large-N tcgen05 backward vs the SIMT kernel for a few configurations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_edgewise import _rand_problem, _run_gpu
from gpu_util import bf16_round, rel_to_max
import mop_b200.functional as MF

zero_ws = "--zero" in sys.argv
if zero_ws:
    _empty = torch.empty
    def empty(*a, **k):
        t = _empty(*a, **k)
        if k.get("dtype") == torch.uint8:
            t.zero_()
        return t
    MF.torch.empty = empty

for (B, H, N, dk, V, r) in [(1, 2, 196, 64, 5, 4), (1, 2, 196, 64, 5, 3), (1, 2, 196, 64, 4, 4), (1, 2, 196, 32, 5, 4), (1, 1, 196, 64, 5, 4), (1, 2, 192, 64, 5, 4), (1, 2, 128, 64, 5, 4)]:
    qkv, scales, head, logit, dy = _rand_problem(B, H, N, dk, V, True, "lowrank", False, r, seed=11 * N + V)
    qkv, dy = bf16_round(qkv), bf16_round(dy)
    _, g_tc = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="tcgen05")
    _, g_s = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="simt")
    torch.cuda.synchronize()
    print((B, H, N, dk, V, r), "zero_ws" if zero_ws else "", {k: round(rel_to_max(g_tc[k].reshape(v.shape), v), 4) for k, v in g_s.items()}, flush=True)
