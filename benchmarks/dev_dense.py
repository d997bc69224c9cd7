"""dev: time / profile the dense(+k3) gate-head Edgewise path (fp32-math kernels) at the config-2 core shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mop_b200
from mop_b200 import functional as MF
B = int(sys.argv[1]) if len(sys.argv) > 1 else 74
H, N, dk, V = 4, 64, 56, 5
C = 2 * V + 2
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
qkv = rn(B, N, 1, 3, H, dk).bfloat16().requires_grad_(True)
sc = [(1 + 0.1 * rn(V, H, 1, dk)).requires_grad_(True) for _ in range(3)]
head = {"conv1.weight": (rn(16, C, 1, 1) / math.sqrt(C)).requires_grad_(True), "conv1.bias": (0.2 * rn(16)).requires_grad_(True),
        "mid3.weight": (rn(16, 16, 3, 3) / 12).requires_grad_(True), "mid3.bias": (0.2 * rn(16)).requires_grad_(True),
        "conv2.weight": (rn(4, 16, 1, 1) / 4).requires_grad_(True), "conv2.bias": (0.3 * rn(4)).requires_grad_(True)}
lg = torch.tensor(-2.0, device="cuda", requires_grad=True)
dy = rn(B, N, H, dk).bfloat16()
MF.kernel_timing = True
for _ in range(3):
    y = mop_b200.edgewise_attention(qkv, *sc, lg, head, n_views=V, beta_not=0.5, gate_mode="dense", use_k3=True)
    y.backward(dy)
torch.cuda.synchronize()
for k, v in MF.kernel_events.items():
    print(k, [round(a.elapsed_time(b), 3) for a, b in v])
