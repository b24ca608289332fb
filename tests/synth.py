"""Seeded synthetic events and configs shared by the golden generator, the tests and bench.py.

Shapes follow the reference batch layout after ``Trainer._prepare_batch`` (scripts/train.py:468-473):
frames / masked_frames / masks are [B, T, 1, H, W] float32, masks in {0,1}, masked = frames * mask
(p2igan_bench/data/sti_dataset.py:223-224).
"""
from __future__ import annotations

import torch


def make_cfg(h: int = 128, w: int = 128, sample_length: int = 16) -> dict:
    """The subset of p2igan_gan_baseline.json the hot path actually reads."""
    return {
        "seed": 2024,
        "model": {"name": "p2igan", "in_channels": 1, "out_channels": 1, "base_channels": 64},
        "data": {"train": {"w": w, "h": h, "sample_length": sample_length, "mask": {"type": "stis", "keep": 4}}},
        "loss": {"adversarial_weight": 0.01, "k1_weight": 0.05, "gan_loss": "hinge", "use_gan": 1},
        "train": {"optimizer": {"type": "Adam", "beta1": 0.0, "beta2": 0.99, "lr": 1e-4}, "batch_size": 16},
    }


def make_mask(B: int, T: int, H: int, W: int, n_obs: int, seed: int, tie_free: bool = False) -> torch.Tensor:
    """[B,T,1,H,W] 0/1 mask.  Default: ONE spatial pattern of ``n_obs`` pixels repeated over T and B
    (the reference's 'stis' gauge mask).  tie_free: an independent pattern per frame and sample, which
    removes the systematic t-1/t+1 equidistance ties."""
    g = torch.Generator().manual_seed(seed)
    if not tie_free:
        m = torch.zeros(H * W)
        m[torch.randperm(H * W, generator=g)[:n_obs]] = 1
        return m.reshape(1, 1, 1, H, W).expand(B, T, 1, H, W).contiguous()
    m = torch.zeros(B * T, H * W)
    for i in range(B * T):
        m[i, torch.randperm(H * W, generator=g)[:n_obs]] = 1
    return m.reshape(B, T, 1, H, W).contiguous()


def make_batch(B: int, T: int, H: int, W: int, n_obs: int, seed: int, tie_free: bool = False):
    """frames = u^4 with u~U(0,1) (sparse-rain-like, mean 0.2); returns (frames, masked, masks)."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.rand(B, T, 1, H, W, generator=g) ** 4
    masks = make_mask(B, T, H, W, n_obs, seed + 7919, tie_free)
    return frames, frames * masks, masks
