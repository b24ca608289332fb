"""p2igan_bench.metrics -> p2igan_b200 (reference: p2igan_bench/metrics/metric.py)."""
from p2igan_b200 import MetricConfig, RainfallMetricSuite, transform  # noqa: F401

__all__ = ["MetricConfig", "RainfallMetricSuite", "transform"]
