"""Device-side batch preparation.

The reference builds every sample on the host (``STIDataset.post_process``, p2igan_bench/data/sti_dataset.py:203-229:
``uint8 -> float32 / 255``, mask multiply, centre crop) and ships three float32 tensors per batch which
``Trainer._prepare_batch`` (scripts/train.py:468-473) permutes to ``[B, T, 1, H, W]`` and copies to the device:
12 bytes per pixel over PCIe.  ``prepare_batch_u8`` takes the raw ``uint8`` frames (1 byte per pixel) and the
gauge mask, and produces the same three tensors with one kernel on the device.
"""
from __future__ import annotations

from typing import Tuple

import torch

from ._lib import LIB, ptr, require_cuda, stream


def prepare_batch(batch) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``Trainer._prepare_batch``: (frames, masked, masks) as [B, T, H, W, 1] float32 (the DataLoader's layout)
    -> [B, T, 1, H, W] views on the current CUDA device."""
    dev = torch.device("cuda", torch.cuda.current_device())
    return tuple(t.permute(0, 1, 4, 2, 3).to(dev, non_blocking=True) for t in batch)


def prepare_batch_u8(frames_u8: torch.Tensor, mask_u8: torch.Tensor, height: int, width: int):
    """frames_u8 [B, T, H0, W0] uint8 (CUDA), mask_u8 [H0, W0] | [B, H0, W0] | [B, T, H0, W0] uint8 (non-zero =
    observed) -> (frames, masked_frames, masks), each [B, T, 1, height, width] float32, centre-cropped."""
    require_cuda(frames_u8, mask_u8)
    if frames_u8.dtype != torch.uint8 or mask_u8.dtype != torch.uint8:
        raise ValueError("prepare_batch_u8 expects uint8 frames and mask")
    if frames_u8.dim() == 5 and frames_u8.shape[-1] == 1:
        frames_u8 = frames_u8[..., 0]
    if frames_u8.dim() != 4:
        raise ValueError(f"frames_u8 must be [B, T, H0, W0], got {tuple(frames_u8.shape)}")
    B, T, H0, W0 = frames_u8.shape
    if tuple(mask_u8.shape) == (H0, W0):
        mode = 0
    elif tuple(mask_u8.shape) == (B, H0, W0):
        mode = 1
    elif tuple(mask_u8.shape) == (B, T, H0, W0):
        mode = 2
    else:
        raise ValueError(f"mask shape {tuple(mask_u8.shape)} does not match frames {tuple(frames_u8.shape)}")
    frames_u8, mask_u8 = frames_u8.contiguous(), mask_u8.contiguous()
    out = [torch.empty(B, T, 1, height, width, dtype=torch.float32, device=frames_u8.device) for _ in range(3)]
    LIB.call("p2i_batch_prep_u8", ptr(frames_u8), ptr(mask_u8), ptr(out[0]), ptr(out[1]), ptr(out[2]), B, T, H0, W0, height, width,
             mode, stream())
    return tuple(out)
