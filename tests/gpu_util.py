"""Shared helpers for the -m gpu parity tests (they call the CUDA path through the C ABI)."""
import torch


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def rel_to_max(a, b):
    """(max abs error) / (max abs reference value): the bf16-mode metric of SURVEY.md 8c."""
    ref = b.detach().double().cpu()
    return max_abs(a, b) / max(ref.abs().max().item(), 1e-30)


def scaled_tol(ref, tol):
    """fp32-mode bound: tol absolute for O(1) tensors, tol * max|ref| for larger ones."""
    return tol * max(1.0, ref.detach().abs().max().item())


def bf16_round(t):
    return t.to(torch.bfloat16).to(t.dtype)
