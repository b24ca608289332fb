"""In-memory stand-ins for the reference's uninstallable dependencies (torchmetrics, mlflow, zarr, h5py, decord,
matplotlib: not in this image nor in /opt/wheelhouse) so that the UNMODIFIED reference checkout can be
imported in the build container: by the golden generators, by the drop-in tests and by bench.py's reference arm.
Test infrastructure only -- nothing here computes anything the parity tests compare."""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("P2I_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "p2igan_bench", "__init__.py"))


def install_stubs() -> None:
    import torch

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import torchmetrics  # noqa: F401
    except ImportError:
        class Metric(torch.nn.Module):
            """torchmetrics.Metric as far as metric.py / losses.py use it: sum-reduced states registered as buffers."""

            def __init__(self):
                super().__init__()
                self._defaults = {}

            def add_state(self, name, default, dist_reduce_fx=None):
                self._defaults[name] = default.clone()
                self.register_buffer(name, default.clone())

            def reset(self):
                for k, v in self._defaults.items():
                    setattr(self, k, v.clone())

        class StructuralSimilarityIndexMeasure(Metric):
            def __init__(self, data_range=1.0, **kw):
                super().__init__()

            def update(self, *a):
                pass

            def compute(self):
                return torch.tensor(float("nan"))

        stub("torchmetrics", Metric=Metric)
        stub("torchmetrics.image", StructuralSimilarityIndexMeasure=StructuralSimilarityIndexMeasure)
    for name in ("mlflow", "zarr", "h5py", "decord", "matplotlib", "matplotlib.cm"):
        try:
            __import__(name)
        except ImportError:
            stub(name)
    if not hasattr(sys.modules["decord"], "VideoReader"):
        sys.modules["decord"].VideoReader = object
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "cm"):
        sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
