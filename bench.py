#!/usr/bin/env python
"""Benchmark of the MoP attention hot path: ViT-MoP (model "E+") training throughput.

Workload (BASELINE.json configs[1]): ViTEdgewise E+ - dim 224, depth 8, heads 4,
100 classes, mlp_ratio 3, 5 views, share_qkv, use_k3, lowrank:mix5 gates (rank 4) -
on synthetic CIFAR-shaped data (32x32, batch 256 per GPU), bf16 autocast,
forward + cross-entropy + backward + AdamW step.  One "step" = one such pass over
one batch.  Batch-sharded data parallel (DDP/NCCL) for --gpus N > 1 (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo (CUDA kernels)
    python bench.py --impl reference ...                       # CPU port of the reference path

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM
(device-timed, CUDA events per step, L2 flushed between steps, max over ranks);
`e2e` = the same step driven from pinned HOST buffers through the public module
API (H2D of the batch + D2H of the loss inside the timed region);
`roofline` = achieved algorithmic TFLOP/s of the dominant attention kernel
against the measured dense-bf16 peak; `cpu_baseline` = the oracle port of the
reference's eager path timed on this box's host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

MODEL = dict(dim=224, depth=8, heads=4, n_classes=100, mlp_ratio=3.0, n_views=5, share_qkv=True, use_k3=True,
             gate_mode="lowrank", gate_rank=4, gate_init="mix5", drop_path=0.1)
BATCH, IMG, PATCH, NTOK = 256, 32, 4, 64
CONFIG = "vit_e_cifar"
# BASELINE.json configs[2]: ViT-B/16-MoP, 224x224 (196 tokens), 86,624,652 parameters, batch 256 per GPU
VIT_B16 = dict(dim=768, depth=12, heads=12, n_classes=1000, mlp_ratio=4.0, n_views=5, share_qkv=True, use_k3=True,
               gate_mode="lowrank", gate_rank=4, gate_init="mix5", drop_path=0.4)


def select_config(name: str, batch: int = 0):
    """`vit_e_cifar` (default, BASELINE.json configs[1], the headline) or `vit_b16` (configs[2], the 8-GPU model)."""
    global MODEL, BATCH, IMG, PATCH, NTOK, WORKLOAD, CONFIG
    CONFIG = name
    if name == "vit_b16":
        MODEL, IMG, PATCH, NTOK = VIT_B16, 224, 16, 196
        BATCH = batch or 256
        WORKLOAD = (f"ViTEdgewise ViT-B/16-MoP (dim768 depth12 heads12 V5 share_qkv use_k3 lowrank:mix5 r4, 86.6 M parameters), "
                    f"ImageNet-shaped 224x224 (196 tokens), batch {BATCH}/GPU, fwd+bwd+AdamW")
    elif batch:
        BATCH = batch
        WORKLOAD = WORKLOAD.replace("batch 256/GPU", f"batch {BATCH}/GPU")
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (ew64::edgewise_bwd3_kernel, B*H = 1024
# problems) from the committed `ncu --set full` capture (a citation, not measured in this run): algorithmic read bytes are
# 29.4 MB (Q, K, V, dy) + 17.8 MB (the forward's row statistics / feature means / gate factors) = 47.2 MB, i.e. no re-reads
# (the 22 MB of dqkv written stay in L2 until after the launch).
DOMINANT_KERNEL_DRAM_BYTES = 47314688 + 1023488
DOMINANT_KERNEL_DRAM_SOURCE = "profiles/r02_ew64_ncu_full_raw.csv (ncu --set full, one launch of edgewise_bwd3_kernel; cited constant)"
WORKLOAD = "ViTEdgewise E+ (dim224 depth8 heads4 V5 share_qkv use_k3 lowrank:mix5 r4), CIFAR-shaped 32x32, batch 256/GPU, fwd+bwd+AdamW"


def edgewise_fwd_flops(B, H, N, dk, V, r):
    """Algorithmic FLOPs of one Edgewise forward (dense contractions only)."""
    return B * H * (2 * N * N * dk * V + 2 * N ** 3 * 2 * (V - 1) + 2 * N * N * dk * 2 + 2 * N * N * 4 * r)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return dict(bf16=float(d["bf16_tflops_sustained"]), bf16_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    except Exception:
        return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle sampling during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's eager path, on the host cores
# ----------------------------------------------------------------------------------------------
def _cpu_step_fn(sample_batch: int):
    """One eager fp32 fwd + CE + bwd + AdamW step of the bench model on the host cores.  Returns (run, kind).

    kind "reference": the UNMODIFIED reference modules from baseline/_ref (baseline/ref_models.py); kind "port": the
    oracle restatement (oracle/vit_edgewise_ref.py) when the reference package is not staged on this machine."""
    import mop_b200
    torch.manual_seed(0)
    x = torch.randn(sample_batch, 3, IMG, IMG)
    y = torch.randint(0, MODEL["n_classes"], (sample_batch,))
    try:
        from baseline.ref_models import reference_vit_edgewise
        model = reference_vit_edgewise(num_tokens=NTOK, patch=PATCH, **MODEL)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)

        def run():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
            return float(loss.detach())
        return run, "reference"
    except ImportError:
        pass
    from oracle.edgewise import EdgewiseConfig
    from oracle.vit_edgewise_ref import train_step_cpu
    skeleton = mop_b200.ViTEdgewise(num_tokens=NTOK, patch=PATCH, compat_experiments_init=False, **MODEL)  # parameters only (same init as the GPU arm)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in skeleton.state_dict().items()}
    cfg = EdgewiseConfig(dim=MODEL["dim"], heads=MODEL["heads"], n_views=MODEL["n_views"], share_qkv=True, use_k3=True,
                         gate_mode="lowrank", gate_rank=MODEL["gate_rank"], gate_init="mix5")
    opt = torch.optim.AdamW([p for p in sd.values() if p.requires_grad], lr=1e-3, weight_decay=0.05)
    return (lambda: train_step_cpu(sd, cfg, x, y, depth=MODEL["depth"], patch=PATCH, drop_path_rate=MODEL["drop_path"], opt=opt)), "port"


def cpu_throughput(steps: int, warmup: int, sample_batch: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind = _cpu_step_fn(sample_batch)
    for _ in range(warmup):
        run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    mean_s = sum(times) / len(times)
    what = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port of the reference eager path"
    return dict(value=sample_batch / mean_s, unit="images/s", cores=cores, kind=kind,
                sample=f"{steps} eager fp32 fwd+bwd+AdamW steps of the same model at batch {sample_batch} (of {BATCH}) after {warmup} "
                       f"warm-up, mean step, {what}, torch CPU {cores} threads"), mean_s


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path at the bench config (batch 256), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # ViT-B/16 at batch 256 takes minutes per step on host cores: a bounded sample of the same workload (8 images per step)
    sample = BATCH if CONFIG == "vit_e_cifar" else 8
    cb, mean_s = cpu_throughput(args.steps, args.warmup, sample)
    line = {"impl": "reference", "metric": "vit_mop_train_images_per_sec", "value": cb["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_s * 1e3 * BATCH / sample,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH,
                       "note": "reference eager path on the host cores only (kind: see cpu_baseline); no GPU work"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# this repo
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import mop_b200
    from mop_b200 import functional as MF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1
    if ddp:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = mop_b200.ViTEdgewise(num_tokens=NTOK, patch=PATCH, compat_experiments_init=False, **MODEL).to(dev)
    model.train()
    # N > 1: every p.grad is a view into one flat buffer and the step does ONE all-reduce (mean) after backward, so that
    # the whole step, collective included, is captured in a CUDA graph (mop_b200/ddp.py).  --no-graph: torch DDP, eager.
    # ViT-B/16 (a 300 ms step, 346 MB of gradients): eager launches, gradient buckets all-reduced WHILE the backward runs
    use_graph = not args.no_graph and CONFIG == "vit_e_cifar"
    from mop_b200.ddp import FlatGradAllReduce
    flat = FlatGradAllReduce(model, bucket_mb=0.0 if use_graph else args.bucket_mb, pack=use_graph) if ddp else None
    net = model
    # bf16 compute copies of the Linear weights, refreshed after every optimizer step (mop_b200/mixed.py): the step loses the
    # ~70 weight-cast / gradient-cast kernels autocast would launch; --no-shadow keeps plain autocast
    from mop_b200.mixed import Bf16Shadow
    shadow = None if args.no_shadow else Bf16Shadow(model)
    # graph mode: parameters re-homed into one flat buffer and fused AdamW over ONE tensor (mop_b200/mixed.py: one kernel instead
    # of four multi-tensor launches over ~150 small tensors; same update).  --no-flat-opt: torch's per-parameter fused AdamW
    flatp = None
    if not args.no_flat_opt and use_graph:
        from mop_b200.mixed import FlatParams
        flatp = FlatParams(model)
        if flat is not None:
            flatp.use_grad_buffer(flat.flat)
    opt = torch.optim.AdamW([flatp.param] if flatp is not None else model.parameters(), lr=1e-3, weight_decay=0.05, fused=True, capturable=use_graph)
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
    x_host = torch.randn(BATCH, 3, IMG, IMG, generator=gen).pin_memory()
    y_host = torch.randint(0, MODEL["n_classes"], (BATCH,), generator=gen).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def fwd_bwd(x, y):
        if flat is not None:
            flat.zero()
        elif flatp is not None:
            flatp.begin()
        else:
            opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(net(x), y)
        loss.backward()
        if flatp is not None and flat is None:
            flatp.pack_grads()
        return loss

    def opt_step():
        opt.step()
        if shadow is not None:
            shadow.refresh()

    def step(x, y):
        loss = fwd_bwd(x, y)
        if flat is not None:
            flat.reduce()
        opt_step()
        return loss

    # ---- the step captured in CUDA graphs and replayed.  N = 1: one graph (fwd + loss + bwd + AdamW).  N > 1: graph A
    #      (fwd + loss + bwd into the flat gradient buffer), ONE eager NCCL all-reduce, graph B (scale + AdamW).
    graph = graph_b = None
    if use_graph:
        static_x, static_y = x_dev.clone(), y_dev.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step(static_x, static_y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        if flat is None:
            if flatp is None:
                opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                static_loss = step(static_x, static_y)
        else:
            with torch.cuda.graph(graph):
                static_loss = fwd_bwd(static_x, static_y)
                if flat.packed:
                    flat.pack()    # the gradient tensors of the captured backward -> the flat buffer (one concatenation)
            if flat.packed:
                flat.bind()        # the optimizer graph reads the reduced gradients from the flat buffer
            graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_b):
                flat.scale()
                opt_step()
        eager_step = step

        def step(x, y):  # noqa: F811  (same signature; inputs are copied into the graph's static buffers)
            if x.data_ptr() != static_x.data_ptr():
                static_x.copy_(x, non_blocking=True)
                static_y.copy_(y, non_blocking=True)
            graph.replay()
            if graph_b is not None:
                flat.all_reduce_sum()
                graph_b.replay()
            return static_loss
        x_dev, y_dev = static_x, static_y

    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    torch.cuda.synchronize()

    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps -------------
    sampler = ClockSampler(local)
    MF.kernel_events.clear()
    launches_per_step = None
    if graph is not None:
        # per-kernel CUDA events cannot live inside a replayed graph: time the attention kernels in 3 eager steps
        # of the same model (same launches, same stream) right before the timed region; the same steps count this
        # repo's kernel launches per step (every ABI call of this model enqueues exactly one kernel)
        MF.kernel_timing = True
        c0 = dict(MF.abi_calls)
        for _ in range(3):
            flush.zero_()
            eager_step(static_x, static_y)
        torch.cuda.synchronize()
        MF.kernel_timing = False
        launches_per_step = sum(MF.abi_calls[k] - c0.get(k, 0) for k in MF.abi_calls) // 3
    else:
        MF.kernel_timing = True
    calls0 = dict(MF.abi_calls)
    if ddp:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(x_dev, y_dev)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if ddp:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    MF.kernel_timing = False
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    kern_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in MF.kernel_events.items()}
    launches = sum(MF.abi_calls[k] - calls0.get(k, 0) for k in MF.abi_calls)
    if graph is not None:   # replayed: the launches of one step (counted above, eagerly) times the timed steps
        launches = launches_per_step * args.steps
    impl_used = dict(MF.last_impl)

    # ---- end to end: host buffers in, loss out, wall clock ------------------------------------
    for _ in range(2):
        float(step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)).detach())
    if ddp:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))
        loss_val = float(loss.detach())  # D2H read of the step's result
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps

    t = torch.tensor([step_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if ddp:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms = t.tolist()

    if rank == 0:
        pk = peaks()
        B, H, N, dk, V, r = BATCH, MODEL["heads"], NTOK, MODEL["dim"] // MODEL["heads"], MODEL["n_views"], MODEL["gate_rank"]
        f_fwd = edgewise_fwd_flops(B, H, N, dk, V, r)
        dom = "edgewise_bwd" if kern_ms.get("edgewise_bwd", 0) >= kern_ms.get("edgewise_fwd", 0) else "edgewise_fwd"
        dom_flops = 2 * f_fwd if dom == "edgewise_bwd" else f_fwd
        dom_ms = kern_ms.get(dom, float("nan"))
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12
        attn_ms = MODEL["depth"] * (kern_ms.get("edgewise_fwd", 0) + kern_ms.get("edgewise_bwd", 0))
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_throughput(steps=4, warmup=1, sample_batch=64) if CONFIG == "vit_e_cifar" else cpu_throughput(steps=1, warmup=0, sample_batch=4)
        extras = {}
        if world == 1 and not args.no_extras:
            # secondary measurements run in a CHILD process after the headline numbers are final: a fault there cannot touch them
            del flush
            torch.cuda.empty_cache()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-only"], capture_output=True, text=True, timeout=420)
                extras = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            except Exception as e:
                extras = {"extras_error": repr(e)[:300]}
        line = {
            "metric": "vit_mop_train_images_per_sec", "value": world * BATCH / (step_ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "name": CONFIG, "global_batch": world * BATCH, "parallelism": f"dp{world}",
                       "grad_allreduce": (None if not ddp else "one flat 16 MB all-reduce between two CUDA graphs" if graph is not None else
                                          f"{len(flat.buckets)} buckets of ~{args.bucket_mb:g} MB launched from grad hooks during the backward"),
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step events)",
                       "attention_impl": impl_used, "loss": loss_val,
                       "precision": ("bf16 autocast, fp32 master weights" + ("" if shadow is None else " with bf16 compute copies of the Linear weights (mop_b200/mixed.py)")),
                       "optimizer": "AdamW lr 1e-3 wd 0.05, fused" + (", one flat parameter buffer" if flatp is not None else ""),
                       "step_launch": "cuda_graph_replay" if graph is not None else "eager"},
            "clocks": clocks,
            "e2e": {"value": world * BATCH / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16"],
                         "traffic": DOMINANT_KERNEL_DRAM_BYTES if CONFIG == "vit_e_cifar" else None,
                         "traffic_source": DOMINANT_KERNEL_DRAM_SOURCE if CONFIG == "vit_e_cifar" else None,
                         "peak_source": pk["source"],
                         "ms_per_launch": dom_ms, "flops_per_launch": dom_flops,
                         "fwd_ms": kern_ms.get("edgewise_fwd"), "bwd_ms": kern_ms.get("edgewise_bwd"),
                         "attention_tflops_fwd_bwd": 3 * f_fwd * MODEL["depth"] / (attn_ms * 1e-3) / 1e12 if attn_ms else None,
                         "attention_share_of_step": attn_ms / step_ms if step_ms else None},
            "cpu_baseline": cb,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if ddp:
        dist.destroy_process_group()


def extra_measurements(dev, flush, pk):
    """Secondary measurements of the same run (N = 1 only; all outside the timed region of the headline):
    * `attention`: BASELINE.json's "MoP attention TFLOP/s" for every named shape - attention core through
      the public functional API, CUDA events, L2 flushed, algorithmic FLOPs (bwd = 2 x fwd), fraction of the measured burst
      bf16 peak (kernels timed in isolation);
    * `variant_dense_k3`: the E+ model with `gate_mode="dense", use_k3=True` (the low-rank head ignores use_k3), eager;
    * `gpu_reference_eager`: the UNMODIFIED reference modules (baseline/_ref) run eagerly on this GPU - the informative GPU baseline."""
    import mop_b200
    out = {}
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    try:
        import attn_microbench as mb
        it = 6
        rows = [mb.edgewise_case("M1/M2 Edgewise E core (B=256,H=4,N=64,dk=56,V=5)", 256, 4, 64, 56, 5, 4, None, it, flush),
                mb.edgewise_case("M3 Edgewise ViT-B/16 core (B=256,H=12,N=196,dk=64,V=5)", 256, 12, 196, 64, 5, 4, "tcgen05", 3, flush),
                mb.quartet_case("M4 Quartet GPT-1024 (B=16,H=12)", 16, 12, 1024, 64, "tcgen05", it, flush),
                mb.quartet_case("M4 Quartet GPT-4096 (B=4,H=12)", 4, 12, 4096, 64, "tcgen05", it, flush),
                mb.sdpa_case("M1 MSA model A (B=256,H=4,N=64,dk=56)", 256, 4, 64, 56, False, "tcgen05", it, flush),
                mb.sdpa_case("M5 Whisper encoder self-attention (B=8,H=16,N=1500)", 8, 16, 1500, 64, False, "tcgen05", it, flush)]
        for r in rows:
            r["frac_of_peak_fwd"] = r["fwd_tflops"] / pk["bf16_burst"]
            r["frac_of_peak_fwd_bwd"] = r["fwd_bwd_tflops"] / pk["bf16_burst"]
            for k in ("dtype", "B", "H", "N", "dk", "V", "r", "causal", "op"):
                r.pop(k, None)
        out["attention"] = {"peak_tflops": pk["bf16_burst"], "peak": "measured burst bf16 (kernels timed alone)", "l2": "flushed",
                            "flops": "algorithmic, dense contractions only, bwd = 2 x fwd", "shapes": rows}
    except Exception as e:   # never lose the headline line to a secondary measurement
        out["attention"] = {"error": repr(e)[:200]}

    def time_steps(model, autocast, steps, warm):
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)
        x = torch.randn(BATCH, 3, IMG, IMG, device=dev)
        y = torch.randint(0, MODEL["n_classes"], (BATCH,), device=dev)
        ts = []
        for i in range(warm + steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                loss = F.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
            e1.record()
            torch.cuda.synchronize()
            if i >= warm:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)
    try:
        kw = dict(MODEL, gate_mode="dense")
        torch.manual_seed(0)
        m = mop_b200.ViTEdgewise(num_tokens=NTOK, patch=PATCH, compat_experiments_init=False, **kw).to(dev).train()
        ms = time_steps(m, True, 3, 2)
        out["variant_dense_k3"] = {"model": "E+ with gate_mode=dense, use_k3=True (4,116,840 parameters)", "ms_per_step": ms,
                                   "images_per_sec": BATCH / (ms * 1e-3), "step_launch": "eager",
                                   "attention_impl": "fp32-math kernels on bf16 storage (the dense 3x3 gate head has no tensor-core path)"}
        del m
    except Exception as e:
        out["variant_dense_k3"] = {"error": repr(e)[:200]}
    try:
        from baseline.ref_models import reference_vit_edgewise
        res = {}
        for name, ac in (("bf16_autocast", True), ("fp32_tf32", False)):
            torch.manual_seed(0)
            torch.backends.cuda.matmul.allow_tf32 = True
            m = reference_vit_edgewise(num_tokens=NTOK, patch=PATCH, **MODEL).to(dev).train()
            ms = time_steps(m, ac, 3, 2)
            res[name] = {"ms_per_step": ms, "images_per_sec": BATCH / (ms * 1e-3)}
            del m
        # ... and the dense + k3 variant of the reference (cuDNN convolutions on [B*H, 12, 64, 64] feature stacks): the fp32-math
        # dense gate head of this repo is SLOWER than this, see DESIGN.md section 6
        torch.manual_seed(0)
        m = reference_vit_edgewise(num_tokens=NTOK, patch=PATCH, **dict(MODEL, gate_mode="dense")).to(dev).train()
        ms = time_steps(m, True, 3, 2)
        res["dense_k3_bf16_autocast"] = {"ms_per_step": ms, "images_per_sec": BATCH / (ms * 1e-3)}
        del m
        out["gpu_reference_eager"] = dict(res, note="unmodified reference modules (baseline/_ref) run eagerly on this GPU, same model / batch")
    except Exception as e:
        out["gpu_reference_eager"] = {"unavailable": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the attention-shape table, the dense+k3 variant and the reference-eager-on-GPU arm")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-flat-opt", action="store_true", help="per-parameter fused AdamW instead of one flat parameter buffer")
    ap.add_argument("--no-shadow", action="store_true", help="plain autocast for the Linear layers (no bf16 compute copies of their weights)")
    ap.add_argument("--config", default="vit_e_cifar", choices=["vit_e_cifar", "vit_b16"],
                    help="vit_e_cifar: BASELINE.json configs[1] (headline); vit_b16: configs[2], ViT-B/16-MoP at 224x224")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 256)")
    ap.add_argument("--bucket-mb", type=float, default=32.0, help="gradient bucket size of the overlapped all-reduce (eager multi-GPU steps)")
    ap.add_argument("--extras-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.extras_only:
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        print(json.dumps(extra_measurements(dev, torch.empty(256 << 20, dtype=torch.uint8, device=dev), peaks())), flush=True)
        return
    select_config(args.config, args.batch)
    if args.config != "vit_e_cifar":
        args.no_extras = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
