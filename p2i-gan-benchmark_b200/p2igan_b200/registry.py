"""Model registry (reference: p2igan_bench/models/__init__.py:13-46).  Only the north-star model
``p2igan`` is provided; the DeepKriging / toy baselines of the reference are out of scope."""
from __future__ import annotations

from typing import Any, Dict

import torch.nn as nn


def _name(cfg: Dict[str, Any]) -> str:
    return cfg.get("model", {}).get("name", "simple").lower()


def build_generator(cfg: Dict[str, Any]) -> nn.Module:
    if _name(cfg) != "p2igan":
        raise ValueError(f"p2igan_b200 implements model.name == 'p2igan' only (got '{_name(cfg)}')")
    from .generator import P2IGenerator
    return P2IGenerator(cfg)


def build_discriminator(cfg: Dict[str, Any]) -> nn.Module:
    if _name(cfg) != "p2igan":
        raise ValueError(f"p2igan_b200 implements model.name == 'p2igan' only (got '{_name(cfg)}')")
    from .discriminator import P2IDiscriminator
    data_cfg = cfg.get("data_loader") or cfg.get("data", {}).get("train", {})
    seq_channels = cfg.get("model", {}).get("in_channels", 1) * data_cfg.get("sample_length", 16)
    return P2IDiscriminator(in_channels=seq_channels)
