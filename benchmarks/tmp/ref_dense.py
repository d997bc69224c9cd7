import sys, os, time, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
import bench
from baseline.ref_models import reference_vit_edgewise
dev = torch.device("cuda", 0)
bench.select_config("vit_e_cifar", 0)
kw = dict(bench.MODEL, gate_mode="dense")
for ac in (True, False):
    torch.manual_seed(0)
    m = reference_vit_edgewise(num_tokens=bench.NTOK, patch=bench.PATCH, **kw).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    x = torch.randn(bench.BATCH, 3, bench.IMG, bench.IMG, device=dev); y = torch.randint(0, 100, (bench.BATCH,), device=dev)
    ts = []
    for i in range(5):
        torch.cuda.synchronize(); t0 = time.time()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            loss = F.cross_entropy(m(x), y)
        loss.backward(); opt.step()
        torch.cuda.synchronize(); ts.append(time.time() - t0)
    print("reference eager dense+k3", "bf16" if ac else "fp32", [round(t * 1e3, 1) for t in ts], "peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
    del m, opt
