"""p2igan_bench.models -> p2igan_b200 (reference: p2igan_bench/models/__init__.py)."""
from p2igan_b200 import P2IDiscriminator, P2IGenerator, build_discriminator, build_generator  # noqa: F401

__all__ = ["build_generator", "build_discriminator", "P2IGenerator", "P2IDiscriminator"]
