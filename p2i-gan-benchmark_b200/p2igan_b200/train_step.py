"""One GAN training iteration in the reference trainer's order (scripts/train.py:240-326), on sm_100a kernels.

    preds = G(masked, masks);  loss_g = ReconstructionLoss(k1)(preds, frames)
    if use_gan:  D(fake.detach()), D(real) -> 0.5*(hinge_real + hinge_fake) -> D backward / Adam
                 freeze D -> adv = gan_loss(D(preds), real, is_disc=False) * adversarial_weight -> loss_g += adv
    G backward / Adam;  unfreeze D

Data parallel (new relative to the reference): one process per GPU, each rank holds its own events;
gradients are summed with ONE flat NCCL all-reduce per model and the 1/world_size factor is folded into the fused
Adam kernel.  Loss scalars stay on the device (no host synchronisation inside the step).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from .losses import ReconstructionLoss, gan_loss
from .optim import FusedAdam


class GANTrainStep:
    def __init__(self, cfg: Dict[str, Any], generator, discriminator=None, process_group=None):
        self.cfg = cfg
        self.G, self.D = generator, discriminator
        loss_cfg = cfg.get("loss", {})
        self.use_gan = bool(loss_cfg.get("use_gan", 0)) and discriminator is not None
        self.gan_type = loss_cfg.get("gan_loss", "hinge")
        self.adv_w = loss_cfg.get("adversarial_weight", 0.01)
        self.real_label = loss_cfg.get("target_real_label", 1.0)
        self.fake_label = loss_cfg.get("target_fake_label", 0.0)
        self.rec = ReconstructionLoss(k1_alpha=loss_cfg.get("k1_weight", 0.0))
        oc = cfg["train"]["optimizer"]
        betas = (oc.get("beta1", 0.0), oc.get("beta2", 0.99))
        self.opt_g = FusedAdam([p for p in generator.parameters()], lr=oc["lr"], betas=betas)
        self.opt_d = FusedAdam([p for p in discriminator.parameters()], lr=oc["lr"], betas=betas) if discriminator is not None else None
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1

    # ---- data-parallel gradient exchange: one flat all-reduce(sum) per model
    def _allreduce(self, params):
        if self.world == 1:
            return 1.0
        grads = [p.grad for p in params if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)
        return 1.0 / self.world

    def _gan(self, logits, real, is_disc):
        return gan_loss(logits, real, loss_type=self.gan_type, is_disc=is_disc, target_real_label=self.real_label,
                        target_fake_label=self.fake_label)

    def step(self, frames, masked_frames, masks) -> Dict[str, torch.Tensor]:
        """Returns device scalars {rec, pool, reg, adv, dis, total}; call .item() outside the hot loop."""
        G, D = self.G, self.D
        preds = G(masked_frames, masks)
        loss_g, pool, reg = self.rec.tensors(preds, frames)
        out = {"rec": loss_g.detach(), "pool": pool, "reg": reg}
        if self.use_gan:
            for p in D.parameters():
                p.requires_grad_(True)
            logits_fake = D(preds.detach())
            logits_real = D(frames)
            loss_d = (self._gan(logits_real, True, True) + self._gan(logits_fake, False, True)) * 0.5
            self.opt_d.zero_grad(set_to_none=True)
            loss_d.backward()
            self.opt_d.step(grad_scale=self._allreduce(list(D.parameters())))
            for p in D.parameters():
                p.requires_grad_(False)
            adv = self._gan(D(preds), True, False) * self.adv_w
            loss_g = loss_g + adv
            out["adv"], out["dis"] = adv.detach(), loss_d.detach()
        self.opt_g.zero_grad(set_to_none=True)
        loss_g.backward()
        self.opt_g.step(grad_scale=self._allreduce(list(G.parameters())))
        if self.use_gan:
            for p in D.parameters():
                p.requires_grad_(True)
        out["total"] = loss_g.detach()
        return out
