"""p2igan_bench.metrics.metric -> p2igan_b200.metrics (reference: p2igan_bench/metrics/metric.py:232-239 ``__all__``)."""
from p2igan_b200.metrics import (EPS, CategoricalMetrics, FractionalSkillScoreMetric, MetricConfig, RainfallMetricSuite,  # noqa: F401
                                 RegressionMetrics, transform)

__all__ = ["RegressionMetrics", "CategoricalMetrics", "FractionalSkillScoreMetric", "MetricConfig", "RainfallMetricSuite"]
