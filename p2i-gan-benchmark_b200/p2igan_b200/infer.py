"""Sliding-window inference of one event (reference: scripts/infer.py:188-260), batched on the device.

The reference slides a 16-frame window (overlap 12 -> step 4) over an event at batch size 1 and averages the overlaps on
the host with numpy: 4 generator launches + 4 D2H copies for a 16-frame event.  Here all windows of the event form ONE
generator batch, the overlap average / output scale / clip run in one kernel, and a single tensor goes back to the host.
"""
from __future__ import annotations

import torch

from ._lib import LIB, ptr, require_cuda, stream


def window_starts(length: int, stride: int = 16, overlap: int = 12):
    step = max(1, stride - overlap)
    return list(range(0, length, step)), step


def sliding_window_infer(generator, masked_frames: torch.Tensor, masks: torch.Tensor, stride: int = 16, overlap: int = 12,
                         output_scale: float = 255.0, max_windows_per_batch: int = 64) -> torch.Tensor:
    """masked_frames, masks: [1, L, 1, H, W] (one event, any L >= 1) -> [L, 1, H, W] float32 on the device."""
    require_cuda(masked_frames, masks)
    if masked_frames.shape[0] != 1:
        raise ValueError("sliding_window_infer takes one event at a time ([1, L, 1, H, W]), as scripts/infer.py does")
    L = masked_frames.shape[1]
    H, W = masked_frames.shape[-2:]
    starts, step = window_starts(L, stride, overlap)
    # window w = frames [s, s+stride), tail padded by repeating the last frame (infer.py:219-227)
    idx = torch.tensor([[min(s + i, L - 1) for i in range(stride)] for s in starts], device=masked_frames.device)
    mf = masked_frames[0][idx]          # [n_win, stride, 1, H, W]
    mk = masks[0][idx]
    outs = []
    with torch.no_grad():
        for i in range(0, len(starts), max_windows_per_batch):
            outs.append(generator(mf[i:i + max_windows_per_batch].contiguous(), mk[i:i + max_windows_per_batch].contiguous()))
    preds = torch.cat(outs, 0).reshape(len(starts), stride, H * W).contiguous()
    out = torch.empty(L, 1, H, W, dtype=torch.float32, device=preds.device)
    LIB.call("p2i_window_blend", ptr(preds), ptr(out), L, H * W, stride, step, len(starts), float(output_scale), stream())
    return out
