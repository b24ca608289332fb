"""p2igan_bench.models -> p2igan_b200 for model.name == "p2igan" (reference: p2igan_bench/models/__init__.py:13-58);
``DKGenerator``, ``STDKGenerator``, ``SimpleGenerator``, ``SimpleDiscriminator`` and the builders' other branches are the
reference's own (pure PyTorch, out of the hot-path scope) and resolve lazily to the reference checkout."""
from typing import Any, Dict

import torch.nn as nn

from p2igan_b200 import P2IDiscriminator, P2IGenerator  # noqa: F401
from p2igan_b200 import registry as _registry
from .._fallthrough import extended_path as _extended_path
from .._fallthrough import reference_attr as _reference_attr

__path__ = _extended_path(__path__, "models")

__all__ = ["build_generator", "build_discriminator", "SimpleGenerator", "SimpleDiscriminator", "P2IGenerator",
           "P2IDiscriminator", "DKGenerator", "STDKGenerator"]


def _is_p2igan(cfg: Dict[str, Any]) -> bool:
    return cfg.get("model", {}).get("name", "simple").lower() == "p2igan"


def build_generator(cfg: Dict[str, Any]) -> nn.Module:
    """model.name 'p2igan' -> the sm_100a generator; 'dk' / 'stdk' / anything else -> the reference's builder (:13-31)."""
    if _is_p2igan(cfg):
        return _registry.build_generator(cfg)
    return _reference_attr("models", "build_generator", __name__)(cfg)


def build_discriminator(cfg: Dict[str, Any]) -> nn.Module:
    if _is_p2igan(cfg):
        return _registry.build_discriminator(cfg)
    return _reference_attr("models", "build_discriminator", __name__)(cfg)


def __getattr__(name: str):
    return _reference_attr("models", name, __name__)
