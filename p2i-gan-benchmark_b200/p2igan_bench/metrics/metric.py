from p2igan_b200.metrics import EPS, MetricConfig, RainfallMetricSuite, transform  # noqa: F401
