import os, sys
sys.path.insert(0, "/root/repo")
import torch, mop_b200
q, k, v = (torch.randn(8, 1500, 16, 64, device="cuda", dtype=torch.bfloat16) for _ in range(3))
with torch.no_grad():
    y = mop_b200.sdpa(q, k, v, causal=False, impl="tcgen05")
torch.cuda.synchronize()
