"""dev: chain_value_logit gradient of the N=64 kernels vs the fp64 oracle at many problems."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_edgewise import _rand_problem, _run_gpu, _oracle
from gpu_util import bf16_round
for B, H in ((150, 4), (256, 4), (64, 4), (16, 4)):
    dk, V, r = 56, 5, 4
    qkv, scales, head, logit, dy = _rand_problem(B, H, 64, dk, V, True, "lowrank", False, r, seed=B + dk)
    qkv, dy = bf16_round(qkv), bf16_round(dy)
    _, g_tc = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="tcgen05")
    _, g_s = _run_gpu(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5, False, torch.bfloat16, impl="simt")
    _, g_ref = _oracle(qkv, scales, head, logit, dy, V, "lowrank", r, 0.5)
    print(B, H, "logit ref", float(g_ref["logit"]), "tc", float(g_tc["logit"]), "simt", float(g_s["logit"]))
    for k in g_ref:
        ref = g_ref[k].double()
        e1 = (g_tc[k].double().cpu().reshape(ref.shape) - ref).abs().max() / ref.abs().max()
        e2 = (g_s[k].double().cpu().reshape(ref.shape) - ref).abs().max() / ref.abs().max()
        print(f"   {k:18s} tc {float(e1):.4f} simt {float(e2):.4f}  |ref|max {float(ref.abs().max()):.3e}")
