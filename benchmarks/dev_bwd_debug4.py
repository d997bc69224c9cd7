import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_edgewise import _rand_problem, _run_gpu, LARGE_CASES
from gpu_util import bf16_round, rel_to_max
cases = [(4, 2, 196, 64, 5, 4, s) for s in (2161, 1, 2, 3)] + [(8, 2, 196, 64, 5, 4, 2161), (8, 1, 100, 56, 3, 2, 1103), (6, 3, 130, 32, 2, 1, 7)]
for (B, H, N, dk, V, r, seed) in cases:
    qkv, scales, head, logit, dy = _rand_problem(B, H, N, dk, V, True, "lowrank", False, r, seed=seed)
    qkv, dy = bf16_round(qkv), bf16_round(dy)
    _, g_tc = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="tcgen05")
    _, g_s = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="simt")
    torch.cuda.synchronize()
    print((B, H, N, dk, V, r, seed), {k: round(rel_to_max(g_tc[k].reshape(v.shape), v), 4) for k, v in g_s.items()}, flush=True)
