"""p2igan_bench.modules -> p2igan_b200 (reference: p2igan_bench/modules/__init__.py)."""
from p2igan_b200 import ReconstructionLoss, gan_loss  # noqa: F401

__all__ = ["ReconstructionLoss", "gan_loss"]
