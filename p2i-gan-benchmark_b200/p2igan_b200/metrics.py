"""Evaluation metrics on sm_100a (reference: p2igan_bench/metrics/metric.py).

``RainfallMetricSuite(MetricConfig()).to(device)``, ``.update(preds, target)``, ``.compute() -> Dict[str, float]``,
``.reset()`` keep the reference API and key names; ``RegressionMetrics``, ``CategoricalMetrics`` and
``FractionalSkillScoreMetric`` (metric.py:232-239 ``__all__``) are views of the same kernels.  One fused kernel pass per
update (per four thresholds) accumulates MAE/RMSE sums, the contingency tables and the FSS sums of the default box sizes
{1,2,4,8}; other box sizes (1..32) take one extra launch per (threshold, size); states are sum-reduced
(``dist_reduce_fx="sum"`` in the reference), so multi-GPU evaluation shards events and all-reduces ``state`` once.
SSIM follows torchmetrics' published algorithm (gaussian 11x11, sigma 1.5; ``csrc/ssim.cu``); torchmetrics itself is absent
from this image, so that one value is "parity unpinned" (SURVEY.md 8c).  Images smaller than the window report NaN.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist

from ._lib import LIB, ptr, require_cuda, stream

EPS = 1e-10


def transform(output):
    """Normalised value -> rainfall intensity (reference: metric.py:16-20)."""
    if isinstance(output, torch.Tensor):
        return torch.pow(10.0, output * 0.0625) * 0.036
    return (10.0 ** (output * 0.0625)) * 0.036


@dataclass
class MetricConfig:
    thresholds: Sequence[float] = (0.5, 2.0, 4.0, 8.0)
    scales: Sequence[int] = (1, 2, 4, 8)
    apply_transform: bool = True
    data_range: float = 1.0


_FAST_SCALES = (1, 2, 4, 8)          # box sizes of the fused single-pass kernel (the reference's default MetricConfig.scales)
_MAX_SCALE = 32                      # p2i_fss_update's shared-memory staging


class RainfallMetricSuite:
    """MAE / RMSE / SSIM + POD / FAR / CSI / HSS per threshold + FSS per (threshold, box size)  (reference: metric.py:194-229).

    Any number of thresholds and any box sizes in 1..32 are accepted (``MetricConfig`` takes any, metric.py:186-191):
    thresholds are processed four at a time by the fused pass (``p2i_metrics_update``: regression sums, contingency
    tables and the FSS sums of the box sizes {1,2,4,8} in ONE read of the two tensors); other box sizes take one
    ``p2i_fss_update`` launch per (threshold, size)."""

    def __init__(self, config: Optional[MetricConfig] = None, *, _regression: bool = True):
        cfg = config or MetricConfig()
        self.cfg = cfg
        self.thr = [float(t) for t in cfg.thresholds]
        self.scales = [int(s) for s in cfg.scales]
        if any(s < 1 or s > _MAX_SCALE for s in self.scales):
            raise ValueError(f"RainfallMetricSuite: FSS box sizes must lie in 1..{_MAX_SCALE} (got {self.scales})")
        self._regression = _regression
        self._fast = [s for s in _FAST_SCALES if s in self.scales]              # <= 4 by construction
        self._generic = [s for s in dict.fromkeys(self.scales) if s not in _FAST_SCALES]
        thr = self.thr or [0.0]                                                  # the fused pass needs >= 1 threshold
        self._chunks = [thr[i:i + 4] for i in range(0, len(thr), 4)]
        self._thr_c = [(ctypes.c_float * len(c))(*c) for c in self._chunks]
        fast = self._fast or [1]
        self._sc_c = (ctypes.c_int * len(fast))(*fast)
        self._n_fast = len(fast)
        self.device: Optional[torch.device] = None
        self.state = None
        self.scratch = None

    # state layout (float32): one 51-float block per chunk of <= 4 thresholds (p2i_metrics_update's layout), then
    # {score_sum, counts} per (threshold, generic box size)
    def _gen_off(self, ti: int, gi: int) -> int:
        return 51 * len(self._chunks) + 2 * (ti * len(self._generic) + gi)

    def to(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RainfallMetricSuite runs on CUDA (sm_100a) only; there is no CPU path")
        n = 51 * len(self._chunks) + 2 * len(self.thr) * len(self._generic)
        self.state = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.scratch = torch.zeros(50 * len(self._chunks) + 2, dtype=torch.float64, device=self.device)
        self.ssim_state = torch.zeros(2, dtype=torch.float64, device=self.device)      # sum of per-image SSIM, images
        return self

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        require_cuda(preds, target)
        if self.state is None:
            self.to(preds.device)
        if preds.shape != target.shape or preds.dim() < 2:
            raise ValueError("preds and target must have the same shape [..., H, W]")
        H, W = preds.shape[-2], preds.shape[-1]
        p = preds.detach().to(torch.float32).contiguous()
        t = target.detach().to(torch.float32).contiguous()
        N = p.numel() // (H * W)
        if N > 65535:
            raise ValueError("RainfallMetricSuite.update: at most 65535 frames per call (FSS is a per-call mean, metric.py:160-169)")
        for c, thr_c in enumerate(self._thr_c):
            LIB.call("p2i_metrics_update", ptr(p), ptr(t), N, H, W, thr_c, len(self._chunks[c]), self._sc_c, self._n_fast,
                     1 if self.cfg.apply_transform else 0, ptr(self.scratch[50 * c:]), ptr(self.state[51 * c:]), stream())
        for ti, thr in enumerate(self.thr):
            for gi, sc in enumerate(self._generic):
                LIB.call("p2i_fss_update", ptr(p), ptr(t), N, H, W, float(thr), sc, ptr(self.scratch[50 * len(self._chunks):]),
                         ptr(self.state[self._gen_off(ti, gi):]), stream())
        if self._regression and H > 10 and W > 10:
            LIB.call("p2i_ssim_update", ptr(p), ptr(t), N, H, W, 1 if self.cfg.apply_transform else 0,
                     float(self.cfg.data_range), ptr(self.ssim_state), stream())

    def all_reduce(self, group=None) -> None:
        """Sum the streaming state over data-parallel ranks (the reference declares dist_reduce_fx='sum')."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.state, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.ssim_state, op=dist.ReduceOp.SUM, group=group)

    def _regression_values(self, s) -> Dict[str, float]:
        n = torch.clamp(s[2], min=1.0)
        ss = self.ssim_state.detach().cpu()
        ssim = float(ss[0] / ss[1]) if float(ss[1]) > 0 else float("nan")
        return {"mae": float(s[0] / n), "rmse": float(torch.sqrt(s[1] / n)), "ssim": ssim}

    def _categorical_values(self, s) -> Dict[str, float]:
        out: Dict[str, float] = {}
        for i, thr in enumerate(self.thr):
            b = 51 * (i // 4) + 3 + 4 * (i % 4)
            h, m, f, c = s[b], s[b + 1], s[b + 2], s[b + 3]
            pre = f"cat_thr{thr:.2f}"
            out[f"{pre}/pod"] = float(h / (h + m + EPS))
            out[f"{pre}/far"] = float(f / (h + f + EPS))
            out[f"{pre}/csi"] = float(h / (h + m + f + EPS))
            den = (m + f) * (f + c) + (h + m) * (m + c)
            out[f"{pre}/hss"] = float(2 * (h * c - m * f) / (den + EPS))
        return out

    def _fss_values(self, s) -> Dict[str, float]:
        out: Dict[str, float] = {}
        for i, thr in enumerate(self.thr):
            for sc in self.scales:
                if sc in self._fast:
                    b = 51 * (i // 4) + 4 * (i % 4) + self._fast.index(sc)
                    tot, cnt = s[19 + b], s[35 + b]
                else:
                    o = self._gen_off(i, self._generic.index(sc))
                    tot, cnt = s[o], s[o + 1]
                if cnt == 0:
                    continue
                out[f"fss_thr{thr:.2f}_s{sc}"] = float(tot / cnt)
        return out

    def compute(self) -> Dict[str, float]:
        s = self.state.detach().cpu()          # host synchronisation: compute() is outside the hot loop
        out: Dict[str, float] = {}
        if self._regression:
            out.update(self._regression_values(s))
        out.update(self._categorical_values(s))
        out.update(self._fss_values(s))
        return out

    def reset(self) -> None:
        if self.state is not None:
            self.state.zero_()
            self.scratch.zero_()
            self.ssim_state.zero_()


class _SuiteView:
    """One of the reference's three ``Metric`` classes as a view of the fused suite (same update / compute / reset / to)."""

    def __init__(self, cfg: MetricConfig, regression: bool):
        self._suite = RainfallMetricSuite(cfg, _regression=regression)

    def to(self, device):
        self._suite.to(device)
        return self

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        self._suite.update(preds, target)

    def reset(self) -> None:
        self._suite.reset()

    def __call__(self, preds, target):
        self.update(preds, target)
        return self.compute()


class RegressionMetrics(_SuiteView):
    """MAE / RMSE / SSIM (reference: metric.py:28-74)."""

    def __init__(self, apply_transform: bool = True, data_range: float = 1.0):
        super().__init__(MetricConfig(thresholds=(), scales=(), apply_transform=apply_transform, data_range=data_range), True)

    def compute(self) -> Dict[str, float]:
        return self._suite._regression_values(self._suite.state.detach().cpu())


class CategoricalMetrics(_SuiteView):
    """POD / FAR / CSI / HSS per threshold (reference: metric.py:77-134)."""

    def __init__(self, thresholds: Sequence[float]):
        super().__init__(MetricConfig(thresholds=tuple(thresholds), scales=()), False)

    def compute(self) -> Dict[str, float]:
        return self._suite._categorical_values(self._suite.state.detach().cpu())


class FractionalSkillScoreMetric(_SuiteView):
    """FSS per (threshold, box size) (reference: metric.py:137-183)."""

    def __init__(self, thresholds: Sequence[float], scales: Sequence[int]):
        super().__init__(MetricConfig(thresholds=tuple(thresholds), scales=tuple(scales)), False)

    def compute(self) -> Dict[str, float]:
        return self._suite._fss_values(self._suite.state.detach().cpu())
