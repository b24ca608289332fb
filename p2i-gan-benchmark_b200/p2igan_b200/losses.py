"""Losses of the training step on sm_100a kernels (reference: p2igan_bench/modules/losses.py).

``ReconstructionLoss(k1_alpha)(pred, target, mask=None) -> (loss Tensor, {"pool": float, "reg": float})`` and
``gan_loss(logits, target_is_real, *, loss_type, is_disc, ...)`` keep the reference signatures; forward and
backward are fused reductions in libp2i_sm100a.so (closed-form backward, no autograd graph of small ops).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from ._lib import LIB, ptr, require_cuda, stream


class _RecLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, k1_alpha, temperature):
        require_cuda(pred, target)
        B, T = pred.shape[0], pred.shape[1]
        HW = pred[0, 0].numel()
        p = pred.detach().contiguous().float()
        y = target.detach().contiguous().float()
        sums = torch.zeros(2, dtype=torch.float32, device=p.device)
        lse = torch.empty(B * (T - 1), 2, dtype=torch.float32, device=p.device)
        LIB.call("p2i_rec_loss_fwd", ptr(p), ptr(y), B, T, HW, float(temperature), ptr(sums), ptr(lse), stream())
        pool = sums[0] / float(B * T * HW)
        reg = sums[1] / float(B)
        ctx.save_for_backward(p, y, lse)
        ctx.k1, ctx.temp, ctx.dims = float(k1_alpha), float(temperature), (B, T, HW)
        ctx.mark_non_differentiable(pool, reg)
        return pool + k1_alpha * reg, pool, reg

    @staticmethod
    def backward(ctx, g, _gp, _gr):
        p, y, lse = ctx.saved_tensors
        B, T, HW = ctx.dims
        d = torch.empty_like(p)
        g = g.detach().contiguous().float()
        LIB.call("p2i_rec_loss_bwd", ptr(p), ptr(y), ptr(lse), ptr(g), ctx.k1, ctx.temp, ptr(d), B, T, HW, stream())
        return d, None, None, None


class ReconstructionLoss:
    """Weighted-L1 pixel pool + k1_alpha * KL of temporal-difference softmaxes (reference: losses.py:32-48)."""

    def __init__(self, k1_alpha: float = 0.0):
        self.k1_alpha = k1_alpha

    def tensors(self, prediction, target):
        """(loss, pool, reg) as device tensors -- no host synchronisation."""
        return _RecLossFn.apply(prediction, target, self.k1_alpha, 0.1)

    def __call__(self, prediction, target, mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Dict[str, float]]:
        loss, pool, reg = self.tensors(prediction, target)        # `mask` is ignored, as in the reference (:39)
        return loss, {"pool": float(pool), "reg": float(reg)}


_MODES = {("hinge", True, True): 0, ("hinge", True, False): 1, ("hinge", False, None): 2, "nsgan": 3, "lsgan": 4}


class _GanLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, mode, label):
        require_cuda(logits)
        x = logits.detach().contiguous().float()
        out = torch.zeros((), dtype=torch.float32, device=x.device)
        LIB.call("p2i_gan_loss_fwd", ptr(x), x.numel(), mode, float(label), ptr(out), stream())
        ctx.save_for_backward(x)
        ctx.mode, ctx.label, ctx.shape = mode, float(label), logits.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        d = torch.empty_like(x)
        g = g.detach().contiguous().float()
        LIB.call("p2i_gan_loss_bwd", ptr(x), x.numel(), ctx.mode, ctx.label, ptr(g), 1.0, ptr(d), stream())
        return d.view(ctx.shape), None, None


def gan_loss(logits: torch.Tensor, target_is_real: bool, *, loss_type: str = "nsgan", is_disc: Optional[bool] = False,
             target_real_label: float = 1.0, target_fake_label: float = 0.0) -> torch.Tensor:
    """Reference: losses.py:232-253 (AdversarialLoss.forward :209-226)."""
    if loss_type == "hinge":
        if is_disc is None:
            raise ValueError("`is_disc` must be set when using hinge loss.")
        mode = (0 if target_is_real else 1) if is_disc else 2
        return _GanLossFn.apply(logits, mode, 0.0)
    if loss_type not in ("nsgan", "lsgan"):
        raise ValueError(f"Unsupported GAN loss type: {loss_type}")
    label = target_real_label if target_is_real else target_fake_label
    return _GanLossFn.apply(logits, _MODES[loss_type], label)
