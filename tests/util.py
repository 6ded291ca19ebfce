"""Shared helpers for the parity tests."""
import torch


def cuda_cl(t):
    """CPU NCDHW tensor -> CUDA tensor of the same logical shape with channels-last memory."""
    return t.cuda().contiguous(memory_format=torch.channels_last_3d)


def rel_err(a, b):
    """normwise relative error ||a-b|| / ||b||"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, tol, what=""):
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a.detach().float().cpu()).all(), f"{what}: non-finite values"
    e = max_rel(a, b)
    assert e <= tol, f"{what}: max-abs error relative to max |ref| = {e:.3e} > {tol:.1e}"
