"""p2igan_bench.metrics -> p2igan_b200 (reference: p2igan_bench/metrics/__init__.py, metric.py:232-239)."""
from p2igan_b200.metrics import (CategoricalMetrics, FractionalSkillScoreMetric, MetricConfig, RainfallMetricSuite,  # noqa: F401
                                 RegressionMetrics, transform)
from .._fallthrough import extended_path as _extended_path
from .._fallthrough import reference_attr as _reference_attr

__path__ = _extended_path(__path__, "metrics")

__all__ = ["MetricConfig", "RainfallMetricSuite", "RegressionMetrics", "CategoricalMetrics", "FractionalSkillScoreMetric",
           "transform"]


def __getattr__(name: str):
    return _reference_attr("metrics", name, __name__)
