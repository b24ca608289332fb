"""Training-loop semantics of scripts/train.py around the hot path: epoch / step accounting,
validation loss, checkpoints in the reference's format, and the resume the reference lacks.

What changes relative to ``Trainer`` (scripts/train.py:97-485):
* one iteration = ``GANTrainStep`` (optionally replayed as CUDA graphs); the loss scalars stay on the device and are
  read once per ``log_step`` instead of three ``float()`` synchronisations per step (train.py:312-324);
* ``evaluate_rec_loss`` accumulates on the device and reads one number at the end (train.py:369-382 reads one per batch);
* ``save_checkpoint`` writes exactly the reference's dict (train.py:475-485), so scripts/infer.py loads it unchanged;
  ``load_checkpoint`` restores models, Adam state and counters.
Out of scope here: data modules, MLflow, PNG grids (host-side).
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Optional

import torch

from .losses import ReconstructionLoss
from .registry import build_discriminator, build_generator
from .train_step import GANTrainStep, GraphedDPStep, GraphedStep

LOSS_KEYS = ("rec", "pool", "reg", "adv", "dis", "total")


class Trainer:
    def __init__(self, cfg: Dict[str, Any], device: Optional[torch.device] = None, process_group=None, use_graphs: bool = True,
                 log_fn: Optional[Callable[[int, Dict[str, float]], None]] = None, peer_exchange: bool = True):
        self.cfg = cfg
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(cfg.get("seed", 42))
        self.generator = build_generator(cfg).to(self.device)
        self.use_gan = bool(cfg["loss"].get("use_gan", 0))
        self.discriminator = build_discriminator(cfg).to(self.device) if self.use_gan else None
        self.generator.train()
        if self.discriminator is not None:
            self.discriminator.train()
        # multi-GPU: NVLink peer-memory gradient exchange (one graph per step) unless peer_exchange=False (NCCL, three graphs)
        self.ts = GANTrainStep(cfg, self.generator, self.discriminator, process_group=process_group, peer_exchange=peer_exchange)
        self.opt_g, self.opt_d = self.ts.opt_g, self.ts.opt_d
        self.rec_loss = ReconstructionLoss(k1_alpha=cfg["loss"].get("k1_weight", 0.0))
        train_cfg = cfg.get("train", {})
        self.log_every = int(train_cfg.get("log_step", 100))
        self.max_steps = train_cfg.get("iterations")
        self.max_epochs = train_cfg.get("max_epochs")
        self.global_step = 0
        self.epoch = 0
        self.best_val = float("inf")
        self.use_graphs = use_graphs
        self._graphed = None
        self._graph_key = None
        self._eager_left = self.EAGER_STEPS
        self.log_fn = log_fn
        self._acc = torch.zeros(len(LOSS_KEYS), dtype=torch.float32, device=self.device)   # running sums since the last log
        self._acc_n = 0

    # ------------------------------------------------------------------ one iteration
    EAGER_STEPS = 2      # real training steps launched eagerly before the step is captured (lazy tables, caches)

    def _step(self, frames, masked, masks) -> Dict[str, torch.Tensor]:
        if not self.use_graphs:
            return self.ts.step(frames, masked, masks)
        key = tuple(frames.shape)
        if self._graph_key != key:
            self._graph_key, self._graphed, self._eager_left = key, None, self.EAGER_STEPS
        if self._graphed is None:
            if self._eager_left > 0:
                self._eager_left -= 1
                return self.ts.step(frames, masked, masks)
            # capture executes nothing: the first replay below IS this batch's training step
            world = torch.distributed.get_world_size(self.ts.pg) if torch.distributed.is_initialized() else 1
            if world > 1 and not self.ts.peer_exchange:
                self._graphed = GraphedDPStep(self.ts, (frames, masked, masks), warmup=0)
            else:
                self._graphed = GraphedStep(lambda a, b, c: self.ts.step(a, b, c), (frames, masked, masks), warmup=0)
        return self._graphed(frames, masked, masks)

    def train_epoch(self, batches: Iterable) -> float:
        """One pass over ``batches`` (tuples of [B,T,1,H,W] CUDA tensors).  Returns the mean generator loss."""
        self.epoch += 1
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        n = 0
        for frames, masked, masks in batches:
            out = self._step(frames.contiguous(), masked.contiguous(), masks.contiguous())
            vec = torch.stack([out.get(k, total.new_zeros(())) for k in LOSS_KEYS])
            self._acc += vec
            self._acc_n += 1
            total += vec[5]
            n += 1
            self.global_step += 1
            if self.global_step % self.log_every == 0:
                self.flush_log()
            if self.max_steps is not None and self.global_step >= self.max_steps:
                break
        return float(total) / max(1, n)

    def flush_log(self) -> Dict[str, float]:
        """The only host synchronisation of the training loop: mean losses since the previous call."""
        if self._acc_n == 0:
            return {}
        vals = (self._acc / self._acc_n).tolist()
        self._acc.zero_()
        self._acc_n = 0
        rec = dict(zip(LOSS_KEYS, vals))
        if self.log_fn is not None:
            self.log_fn(self.global_step, rec)
        return rec

    def fit(self, train_batches: Callable[[], Iterable], val_batches: Optional[Callable[[], Iterable]] = None,
            save_dir: Optional[str] = None) -> None:
        """Epoch loop of Trainer.fit (train.py:184-238): train, validate, keep best / latest checkpoints."""
        epochs = self.max_epochs
        if epochs is None:
            epochs = 1 if self.max_steps is None else math.inf
        while self.epoch < epochs:
            self.train_epoch(train_batches())
            if val_batches is not None:
                val = self.evaluate_rec_loss(val_batches())
                if save_dir is not None and val < self.best_val:
                    self.best_val = val
                    self.save_checkpoint(Path(save_dir) / "best.pth", self.epoch)
            if save_dir is not None:
                self.save_checkpoint(Path(save_dir) / "latest.pth", self.epoch)
            if self.max_steps is not None and self.global_step >= self.max_steps:
                break

    # ------------------------------------------------------------------ validation (train.py:369-382)
    def evaluate_rec_loss(self, batches: Optional[Iterable]) -> float:
        if batches is None:
            return 0.0
        was_training = self.generator.training
        self.generator.eval()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        n = 0
        with torch.no_grad():
            for frames, masked, masks in batches:
                preds = self.generator(masked, masks)
                loss, _, _ = self.rec_loss.tensors(preds, frames)
                total += loss
                n += 1
        self.generator.train(was_training)
        return float(total) / max(1, n)

    # ------------------------------------------------------------------ checkpoints (train.py:475-485)
    def save_checkpoint(self, path, epoch: Optional[int] = None) -> None:
        state = {
            "epoch": self.epoch if epoch is None else epoch,
            "global_step": self.global_step,
            "generator": self.generator.state_dict(),
            "optimizer_g": self.opt_g.state_dict(),
        }
        if self.discriminator is not None and self.opt_d is not None:
            state["discriminator"] = self.discriminator.state_dict()
            state["optimizer_d"] = self.opt_d.state_dict()
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        torch.save(state, path)

    def load_checkpoint(self, path) -> None:
        """Resume (absent from the reference): models, both Adam states, epoch and step counters."""
        ckpt = torch.load(path, map_location=self.device, weights_only=True)
        self.generator.load_state_dict(ckpt["generator"])
        self.opt_g.load_state_dict(ckpt["optimizer_g"])
        if self.discriminator is not None and "discriminator" in ckpt:
            self.discriminator.load_state_dict(ckpt["discriminator"])
            self.opt_d.load_state_dict(ckpt["optimizer_d"])
        self.epoch = int(ckpt.get("epoch", 0))
        self.global_step = int(ckpt.get("global_step", 0))
        self._graphed, self._graph_key = None, None     # re-capture after the in-place parameter rewrite
