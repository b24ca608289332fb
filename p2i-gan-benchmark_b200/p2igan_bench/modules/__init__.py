"""p2igan_bench.modules -> p2igan_b200 (reference: p2igan_bench/modules/__init__.py).  ``p2igan_bench.modules.layer`` /
``.deconv_pytorch`` (needed by the reference's DeepKriging models, models/dk.py:7) resolve to the reference's files."""
from p2igan_b200 import ReconstructionLoss, gan_loss  # noqa: F401
from .._fallthrough import extended_path as _extended_path
from .._fallthrough import reference_attr as _reference_attr

__path__ = _extended_path(__path__, "modules")

__all__ = ["ReconstructionLoss", "gan_loss"]


def __getattr__(name: str):
    return _reference_attr("modules", name, __name__)
