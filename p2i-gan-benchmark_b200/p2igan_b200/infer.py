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


def run_inference(generator, events, output_path: str, attrs=None, stride: int = 16, overlap: int = 12, output_scale: float = 255.0,
                  passes: int = 1, overwrite: bool = False, max_windows_per_batch: int = 64):
    """The event loop of scripts/infer.py:194-260 around ``sliding_window_infer``: every event (``events`` yields
    (frames, masked_frames, masks), each [1, L, 1, H, W] on the generator's device -- the reference's test loader has batch
    size 1) becomes dataset ``event_XX`` [L, 1, H, W] float32 of a zarr-v2 group; with ``passes`` > 1 the running mean over
    passes is kept (infer.py:258-260; the generator is deterministic, so further passes only matter for stochastic masks).
    One device->host copy per event.  Returns the list of dataset names."""
    from . import zarr_io
    group = zarr_io.open_group(str(output_path), dict(attrs or {}, passes=int(passes), output_scale=float(output_scale)),
                               overwrite=overwrite)
    names = []
    was_training = generator.training
    generator.eval()
    try:
        for pass_idx in range(max(1, int(passes))):
            for offset, (_, masked, masks) in enumerate(events() if callable(events) else events):
                comp = sliding_window_infer(generator, masked, masks, stride=stride, overlap=overlap, output_scale=output_scale,
                                            max_windows_per_batch=max_windows_per_batch).cpu().numpy()
                name = f"event_{offset + 1:02d}"
                if pass_idx == 0:
                    names.append(name)
                else:
                    cur = zarr_io.read_array(group, name)
                    comp = cur + (comp - cur) / float(pass_idx + 1)
                zarr_io.write_array(group, name, comp)
    finally:
        generator.train(was_training)
    return names
