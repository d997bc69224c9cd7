#!/usr/bin/env python
"""Model-level fwd+bwd of BASELINE configs 4 and 5 and of model B, on one GPU.

The models are the REFERENCE's own classes (baseline/_ref): `create_gpt_quartet` / `create_gpt_mop` at GPT-2-small width,
the `WhisperMoP` encoder on 80 x 3000-frame-shaped mels (1500 audio positions), `ViT_MoP`.  Each is built twice with the same
seed - unpatched (the reference run eagerly on this GPU: the informative GPU baseline) and after `mop_b200.dropin.
patch_reference()` (its attention classes, `MoP2D` and the ViT token gate replaced by this repo's kernels) - and timed with
CUDA events over fwd + loss + bwd under bf16 autocast, L2 flushed between iterations.  One JSON line per model.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def time_it(fn, iters, warm, flush):
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    from baseline import ref_models
    ref_models.import_reference()
    import mop.models as mm
    import mop.models.gpt_mop as gm
    import mop.models.whisper_mop as wm
    from mop.models.quartet_attn_patch import TransformerConfig
    import mop_b200.dropin as dropin
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    vocab = 50257

    def gpt(kind, T, B):
        cfg = TransformerConfig(n_layer=12, n_head=12, n_embd=768, dropout=0.0, block_size=T, bias=True, use_quartet=True)
        build = (lambda: gm.create_gpt_quartet(vocab, cfg)) if kind == "quartet" else (lambda: gm.create_gpt_mop(vocab, cfg))
        idx = torch.randint(0, vocab, (B, T), device=dev)

        def step(m):
            out = m(idx)
            logits = out[0] if isinstance(out, tuple) else out
            loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), idx.reshape(-1))
            loss.backward()
        return build, step, B * T, "tokens/s"

    def whisper(B):
        cfg = wm.WhisperConfig(n_layer_dec=0, dropout=0.0)   # encoder only: 12 layers, 1024 wide, 16 heads, 1500 positions
        mel = torch.randn(B, 1500, 80, device=dev)

        def step(m):
            enc, gates = m.encode(mel)
            (enc.float().square().mean() + gates.float().mean()).backward()
        return (lambda: wm.create_whisper_mop(cfg)), step, B * 1500, "frames/s"

    def vit_b(B):
        x = torch.randn(B, 3, 32, 32, device=dev)
        y = torch.randint(0, 100, (B,), device=dev)

        def step(m):
            F.cross_entropy(m(x).float(), y).backward()
        return (lambda: mm.ViT_MoP(dim=256, depth=6, heads=4, n_classes=100)), step, B, "images/s"

    cases = {
        "gpt_quartet_T1024": lambda: gpt("quartet", 1024, 8),
        "gpt_quartet_T4096": lambda: gpt("quartet", 4096, 2),
        "gpt_mop_T1024": lambda: gpt("mop", 1024, 8),
        "whisper_mop_encoder": lambda: whisper(8),
        "vit_mop_model_b": lambda: vit_b(256),
    }
    for name, mk in cases.items():
        if a.only and a.only not in name:
            continue
        build, step, units, unit = mk()
        row = {"model": name, "unit": unit, "dtype": "bf16 autocast", "l2": "flushed", "timed": "fwd + loss + bwd (no optimizer)"}
        for arm in ("reference_eager", "patched"):
            dropin.unpatch_reference()
            if arm == "patched":
                dropin.patch_reference()
            try:
                torch.manual_seed(0)
                m = build().to(dev).train()

                def fn():
                    for p_ in m.parameters():
                        p_.grad = None
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        step(m)
                ms = time_it(fn, a.iters, 2, flush)
                row[arm] = {"ms": ms, "per_sec": units / (ms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
                del m
            except Exception as e:   # e.g. the eager T=4096 Quartet maps do not fit
                row[arm] = {"error": repr(e)[:160]}
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
        dropin.unpatch_reference()
        if "ms" in row.get("reference_eager", {}) and "ms" in row.get("patched", {}):
            row["speedup_vs_reference_eager"] = row["reference_eager"]["ms"] / row["patched"]["ms"]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
