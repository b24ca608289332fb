"""GPU parity: fused metric kernel vs the CPU oracle suite (itself pinned to the reference golden). Gate: abs 1e-3
(north star), observed ~1e-6."""
import math

import pytest
import torch

import synth
from oracle import p2i_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _compare(ours, ref, tol=1e-3):
    """ref without "ssim": reference golden taken with a stubbed torchmetrics (SSIM parity is unpinned)."""
    assert set(ours) - {"ssim"} == set(ref) - {"ssim"}
    for k, v in ref.items():
        assert abs(ours[k] - v) < tol, (k, ours[k], v)


@pytest.mark.parametrize("shape,scale", [((2, 16, 1, 32, 32), 85.0), ((1, 16, 1, 128, 128), 85.0), ((2, 3, 1, 40, 72), 60.0)])
def test_metric_suite_matches_oracle(shape, scale):
    from p2igan_b200.metrics import MetricConfig, RainfallMetricSuite
    g = torch.Generator().manual_seed(3)
    pred = torch.rand(shape, generator=g) ** 2 * scale
    tgt = torch.rand(shape, generator=g) ** 2 * scale
    ref = O.MetricSuiteOracle(with_ssim=True)
    suite = RainfallMetricSuite(MetricConfig()).to(DEV)
    for a, b in ((pred, tgt), (tgt.flip(-1), tgt)):
        ref.update(a, b)
        suite.update(a.to(DEV), b.to(DEV))
    _compare(suite.compute(), ref.compute())
    suite.reset()
    suite.update(pred.to(DEV), tgt.to(DEV))
    ref2 = O.MetricSuiteOracle(with_ssim=True)
    ref2.update(pred, tgt)
    _compare(suite.compute(), ref2.compute())


def test_metric_suite_reference_golden(golden):
    from p2igan_b200.metrics import MetricConfig, RainfallMetricSuite
    frames, _, _ = synth.make_batch(2, 16, 32, 32, 12, 1)
    pred = golden["g32"]["out"]
    suite = RainfallMetricSuite(MetricConfig()).to(DEV)
    suite.update((pred * 85.0).to(DEV), (frames * 85.0).to(DEV))
    suite.update((frames.flip(1) * 85.0).to(DEV), (frames * 85.0).to(DEV))
    _compare(suite.compute(), golden["metrics32"])


def test_metric_suite_stress_256_and_properties():
    """BASELINE configs[4] shape (8 events of 20x256x256): identical inputs => perfect scores; the contingency table
    of every threshold sums to the pixel count."""
    from p2igan_b200.metrics import MetricConfig, RainfallMetricSuite
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(8, 20, 1, 256, 256, generator=g) ** 3 * 85.0).to(DEV)
    suite = RainfallMetricSuite(MetricConfig()).to(DEV)
    suite.update(x, x)
    m = suite.compute()
    assert m["mae"] == 0.0 and m["rmse"] == 0.0
    assert abs(m["ssim"] - 1.0) < 1e-5              # identical images
    for thr in ("0.50", "2.00", "4.00", "8.00"):
        assert abs(m[f"cat_thr{thr}/pod"] - 1.0) < 1e-6 and m[f"cat_thr{thr}/far"] == 0.0
        for s in (1, 2, 4, 8):
            assert abs(m[f"fss_thr{thr}_s{s}"] - 1.0) < 1e-6
    st = suite.state.cpu()
    for t in range(4):
        assert float(st[3 + 4 * t:7 + 4 * t].sum()) == 8 * 20 * 256 * 256
    with pytest.raises(ValueError):
        RainfallMetricSuite(MetricConfig(scales=(1, 33)))          # box sizes 1..32 (documented in INTEGRATION.md 5)


def test_ssim_matches_oracle_restatement_and_small_images():
    """SSIM vs the oracle's restatement of torchmetrics' algorithm (parity unpinned: torchmetrics is absent), on raw and
    transformed values, ragged sizes; images not larger than the 11x11 window report NaN."""
    from p2igan_b200.metrics import MetricConfig, RainfallMetricSuite
    g = torch.Generator().manual_seed(8)
    for shape, tr, scale in (((3, 4, 1, 37, 53), False, 1.0), ((2, 16, 1, 128, 128), True, 60.0), ((1, 2, 1, 11, 30), False, 1.0)):
        pred = torch.rand(shape, generator=g) * scale
        tgt = (pred + 0.2 * scale * torch.rand(shape, generator=g)).clamp(0, scale)
        suite = RainfallMetricSuite(MetricConfig(apply_transform=tr)).to(DEV)
        suite.update(pred.to(DEV), tgt.to(DEV))
        got = suite.compute()["ssim"]
        H, W = shape[-2:]
        p, t = (O.rain_rate(pred), O.rain_rate(tgt)) if tr else (pred, tgt)
        ref = float(O.ssim_per_image(p.reshape(-1, 1, H, W), t.reshape(-1, 1, H, W)).mean())
        assert abs(got - ref) < 1e-3, (shape, got, ref)
    small = RainfallMetricSuite(MetricConfig()).to(DEV)
    small.update(torch.rand(1, 2, 1, 8, 8).to(DEV), torch.rand(1, 2, 1, 8, 8).to(DEV))
    assert math.isnan(small.compute()["ssim"])


def test_categorical_scores_cross_check_against_exp1_golden():
    """Row N4 cross-check: POD / FAR / CSI of the CUDA metric suite vs the values the reference's OTHER implementation
    (experiments/exp1.categorical_metrics via run_exp1, all pixels, no /3) produced on the same arrays.  HSS is excluded: the
    two reference files define it with different denominators (metric.py:126-127 vs exp1.py:169-172)."""
    import os
    from p2igan_b200.metrics import MetricConfig, RainfallMetricSuite
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = torch.load(os.path.join(root, "tests", "golden", "reference_exp1.pt"), weights_only=False)
    crop = g["crop"]
    T, H, W = g["pred"].shape
    top, left = (H - crop) // 2, (W - crop) // 2
    pred = g["pred"][:, top:top + crop, left:left + crop].contiguous()
    truth = g["truth"][:, top:top + crop, left:left + crop].contiguous()
    suite = RainfallMetricSuite(MetricConfig()).to(DEV)
    suite.update(pred[None, :, None].to(DEV), truth[None, :, None].to(DEV))
    m = suite.compute()
    ref = g["results"]["all_0"]
    for thr, key in ((0.5, "CAT_0.5"), (2.0, "CAT_2"), (4.0, "CAT_4"), (8.0, "CAT_8")):
        for ours_k, ref_k in (("pod", "POD"), ("far", "FAR"), ("csi", "CSI")):
            assert abs(m[f"cat_thr{thr:.2f}/{ours_k}"] - ref[key][ref_k]) < 1e-5, (thr, ours_k)


def test_arbitrary_thresholds_and_scales_and_the_three_metric_classes():
    """MetricConfig takes any thresholds / box sizes in the reference (metric.py:186-191): six thresholds (two fused passes)
    and box sizes outside {1,2,4,8} (generic kernel) vs the oracle; RegressionMetrics / CategoricalMetrics /
    FractionalSkillScoreMetric (metric.py:232-239 __all__) report the same values as the suite."""
    from p2igan_b200.metrics import (CategoricalMetrics, FractionalSkillScoreMetric, MetricConfig, RainfallMetricSuite,
                                     RegressionMetrics)
    thr, scales = (0.1, 0.5, 1.0, 2.0, 4.0, 8.0), (1, 3, 5, 8, 16)
    g = torch.Generator().manual_seed(12)
    shape = (2, 5, 1, 48, 72)
    pred = torch.rand(shape, generator=g) ** 2 * 85.0
    tgt = torch.rand(shape, generator=g) ** 2 * 85.0
    ref = O.MetricSuiteOracle(thresholds=thr, scales=scales, with_ssim=True)
    suite = RainfallMetricSuite(MetricConfig(thresholds=thr, scales=scales)).to(DEV)
    reg, cat, fss = RegressionMetrics().to(DEV), CategoricalMetrics(thr).to(DEV), FractionalSkillScoreMetric(thr, scales).to(DEV)
    for a, b in ((pred, tgt), (tgt.flip(-2), tgt)):
        ref.update(a, b)
        for m in (suite, reg, cat, fss):
            m.update(a.to(DEV), b.to(DEV))
    want, got = ref.compute(), suite.compute()
    assert list(got) == list(want)                      # same keys in the reference's order (regression, categorical, FSS)
    assert len(got) == 3 + 4 * len(thr) + len(thr) * len(scales)
    _compare(got, want)
    parts = {}
    for m in (reg, cat, fss):
        parts.update(m.compute())
    assert list(parts) == list(got)
    for k in got:                                   # separate instances: the double-precision atomics sum in a different order
        assert abs(parts[k] - got[k]) <= 1e-6 * max(1.0, abs(got[k])), (k, parts[k], got[k])
    fss.reset()
    assert fss.compute() == {}                          # counts == 0 -> key omitted (metric.py:178-179)


def test_downsample_duplicate_channels_standalone_forward_backward():
    """DownsampleDuplicateChannels.forward (layer.py:205-214) as a module of its own: values and gradient vs torch."""
    import torch.nn.functional as F
    from p2igan_b200.layers import DownsampleDuplicateChannels
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 24, 40, generator=g)
    m = DownsampleDuplicateChannels(length=16)

    def ref_fn(t):
        b, c, h, w = t.shape
        p = F.max_pool2d(t, 2, 2).view(b * 16, c // 16, h // 2, w // 2)
        return p.repeat_interleave(2, dim=1).view(b, 2 * c, h // 2, w // 2)

    xr = x.clone().requires_grad_(True)
    yr = ref_fn(xr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    xc = x.to(DEV).requires_grad_(True)
    y = m(xc)
    y.backward(gy.to(DEV))
    assert torch.equal(y.detach().cpu(), yr.detach())
    assert torch.allclose(xc.grad.cpu(), xr.grad, atol=1e-6)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 24, 8, 8, device=DEV))
