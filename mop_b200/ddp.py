"""Batch-sharded data parallelism for the MoP models (one process per GPU).

The attention path shards by batch: every (sample, head) problem is independent, so the
only collective is the gradient all-reduce of torch DDP over NCCL (NVLink/NVSwitch).  The
reference has no multi-GPU code; this module is the whole of ours.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process if unset)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group when WORLD_SIZE > 1.  nccl on GPU, gloo on CPU."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def wrap(model: torch.nn.Module, local_rank: int = 0, bucket_cap_mb: int = 64) -> torch.nn.Module:
    """DDP wrapper with a single large bucket for small models (4 M params = 16 MB: one all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    on_gpu = next(model.parameters()).is_cuda
    return torch.nn.parallel.DistributedDataParallel(
        model, device_ids=[local_rank] if on_gpu else None, bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)


def max_over_ranks(value: float, device) -> float:
    """Timing rule: report the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class FlatGradAllReduce:
    """One flat fp32 gradient buffer for the whole model and ONE all-reduce per step.

    torch DDP's reducer hooks are not CUDA-graph friendly and, for the small ViT-MoP models (4 M parameters = 16 MB of
    gradients), bucketing/overlap buys nothing: the step is launch bound.  Here every ``p.grad`` is a view into one flat
    buffer, so ``reduce()`` is a single NCCL all-reduce (mean) between a captured forward + backward graph and a
    captured scale + optimizer graph.  Use ``zero()`` instead of ``optimizer.zero_grad(set_to_none=True)``.
    """

    def __init__(self, model: torch.nn.Module, bucket_mb: float = 0.0, pack: bool = False):
        """``pack=True`` (small, launch-bound models; needs ``bucket_mb == 0``): the gradients are NOT accumulated into views of
        the flat buffer (that costs one ``add_`` kernel per parameter and step: autograd adds into an existing ``.grad``).  Instead
        every step starts with ``begin()`` (``p.grad = None``: autograd keeps the freshly computed gradient tensors), ``pack()``
        gathers them into the flat buffer with one concatenation, and ``bind()`` points every ``p.grad`` at its slice of the
        (reduced) flat buffer for the optimizer.  Under CUDA graphs the gradient tensors of the captured backward have fixed
        addresses, so ``begin`` / ``pack`` are captured with the backward and ``bind`` once before the optimizer graph.

        ``bucket_mb > 0``: overlapped mode for models whose gradients are large (ViT-B/16: 346 MB).  The flat buffer is cut
        into buckets of about that size in REVERSE parameter order (the order the backward produces gradients); a
        post-accumulate-grad hook on every parameter counts its bucket down and, when a bucket is complete, launches its
        all-reduce asynchronously on NCCL's stream while the backward continues.  ``reduce()`` then only waits and scales.
        (Eager steps only: the hooks are not captured by CUDA graphs - small models use ``bucket_mb = 0`` and one all-reduce.)"""
        params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params = params
        self.packed = bool(pack)
        if self.packed and bucket_mb > 0:
            raise ValueError("pack=True is the one-all-reduce mode (bucket_mb must be 0)")
        self.views = []
        off = 0
        for p in params:
            if p.dtype != torch.float32:
                raise TypeError("FlatGradAllReduce expects fp32 parameters (use autocast for bf16 compute)")
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            if not self.packed:
                p.grad = self.views[-1]
            off += n
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:  # same initial weights on every rank
            for p in model.parameters():
                dist.broadcast(p.data, 0)
        # ---- overlapped mode: buckets over the flat buffer, filled from its END (reverse parameter order) ----
        self.buckets = []          # (lo, hi) element ranges of the flat buffer, in launch order
        self._pending = []         # parameters still missing per bucket in the current step
        self._works = []
        self._hooks = []
        if bucket_mb > 0 and self.world > 1:
            cap = int(bucket_mb * (1 << 20) / 4)
            bounds, hi, size, members, off = [], total, 0, [], total
            for p in reversed(params):
                off -= p.numel()
                size += p.numel()
                members.append(p)
                if size >= cap:
                    bounds.append((off, hi, members))
                    hi, size, members = off, 0, []
            if members:
                bounds.append((off, hi, members))
            self.buckets = [(lo, hi_) for lo, hi_, _ in bounds]
            self._counts = [len(m) for _, _, m in bounds]
            self._pending = list(self._counts)
            for bi, (_, _, mem) in enumerate(bounds):
                for p in mem:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi: int):
        def hook(_param):
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                lo, hi = self.buckets[bi]
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        return hook

    # ---- pack mode ----
    def begin(self) -> None:
        """Start of a step: autograd will keep (not add) the gradients it computes."""
        for p in self.params:
            p.grad = None

    def pack(self) -> None:
        """After the backward: gather the gradients into the flat buffer (one concatenation; a missing gradient is zero)."""
        gs = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(gs, out=self.flat)

    def bind(self) -> None:
        """Point every p.grad at its slice of the flat buffer (what the optimizer reads after the all-reduce)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self) -> None:
        if self.packed:
            self.begin()
            return
        self.flat.zero_()
        if self.buckets:
            self._pending = list(self._counts)
            self._works = []

    def all_reduce_sum(self) -> None:
        if self.world <= 1:
            return
        if self.buckets:   # launched bucket by bucket during the backward: wait for them (and catch a bucket no hook completed)
            for bi, left in enumerate(self._pending):
                if left > 0:   # parameters without a gradient this step
                    lo, hi = self.buckets[bi]
                    self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                    self._pending[bi] = 0
            for w in self._works:
                w.wait()
            self._works = []
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def scale(self) -> None:
        if self.world > 1:
            self.flat.mul_(1.0 / self.world)

    def reduce(self) -> None:
        if self.packed:
            self.pack()
        self.all_reduce_sum()
        self.scale()
        if self.packed:
            self.bind()
