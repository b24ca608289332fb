"""p2igan_bench.models.p2igan -> p2igan_b200 (reference: p2igan_bench/models/p2igan.py:23-183)."""
from p2igan_b200.discriminator import P2IDiscriminator  # noqa: F401
from p2igan_b200.generator import P2IGenerator  # noqa: F401
from p2igan_b200.layers import EBlock  # noqa: F401

__all__ = ["P2IGenerator", "P2IDiscriminator", "EBlock"]
