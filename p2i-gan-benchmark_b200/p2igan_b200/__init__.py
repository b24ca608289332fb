"""B200-native (sm_100a) implementation of the P2I-GAN training/inference hot path.

Public surface mirrors what the reference's scripts import:
``build_generator``, ``build_discriminator``, ``P2IGenerator``, ``P2IDiscriminator``,
``ReconstructionLoss``, ``gan_loss``, ``MetricConfig``, ``RainfallMetricSuite``; plus the pieces the reference does
not have: ``FusedAdam``, ``GANTrainStep`` (one iteration in scripts/train.py's order, data parallel, CUDA-graph
capturable), ``sliding_window_infer`` (scripts/infer.py's window loop, batched), ``Trainer`` (scripts/train.py's loop
semantics: validation loss, reference-format checkpoints, resume) and ``prepare_batch_u8`` (device-side batch preparation).
Every op is a hand-written CUDA kernel in ``libp2i_sm100a.so`` (C ABI: include/p2i_b200.h);
there is no CPU, PyTorch-eager or Triton fallback.
"""
from ._overlap import set_stream_overlap  # noqa: F401
from .data import prepare_batch, prepare_batch_u8  # noqa: F401
from .discriminator import P2IDiscriminator  # noqa: F401
from .generator import P2IGenerator  # noqa: F401
from .infer import run_inference, sliding_window_infer  # noqa: F401
from .losses import ReconstructionLoss, gan_loss  # noqa: F401
from .metrics import (CategoricalMetrics, FractionalSkillScoreMetric, MetricConfig, RainfallMetricSuite,  # noqa: F401
                      RegressionMetrics, transform)
from .optim import FusedAdam  # noqa: F401
from .registry import build_discriminator, build_generator  # noqa: F401
from .train_step import FlatGrads, GANTrainStep, GraphedDPStep, GraphedStep  # noqa: F401
from .trainer import Trainer  # noqa: F401

__all__ = ["P2IGenerator", "P2IDiscriminator", "build_generator", "build_discriminator", "ReconstructionLoss", "gan_loss",
           "MetricConfig", "RainfallMetricSuite", "RegressionMetrics", "CategoricalMetrics", "FractionalSkillScoreMetric",
           "transform", "FusedAdam", "GANTrainStep", "GraphedStep", "FlatGrads",
           "GraphedDPStep", "sliding_window_infer", "run_inference", "Trainer", "prepare_batch", "prepare_batch_u8", "set_stream_overlap"]
