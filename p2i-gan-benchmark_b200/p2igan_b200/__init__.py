"""B200-native (sm_100a) implementation of the P2I-GAN training/inference hot path.

Public surface mirrors what the reference's scripts import:
``build_generator``, ``build_discriminator``, ``P2IGenerator``, ``P2IDiscriminator``,
``ReconstructionLoss``, ``gan_loss``, ``MetricConfig``, ``RainfallMetricSuite``.
Every op is a hand-written CUDA kernel in ``libp2i_sm100a.so`` (C ABI: include/p2i_b200.h);
there is no CPU, PyTorch-eager or Triton fallback.
"""
from .generator import P2IGenerator  # noqa: F401
from .registry import build_discriminator, build_generator  # noqa: F401

__all__ = ["P2IGenerator", "build_generator", "build_discriminator"]
