"""Evaluation metrics on sm_100a (reference: p2igan_bench/metrics/metric.py).

``RainfallMetricSuite(MetricConfig()).to(device)``, ``.update(preds, target)``, ``.compute() -> Dict[str, float]``,
``.reset()`` keep the reference API and key names.  One fused kernel pass per update accumulates MAE/RMSE sums, the
contingency tables of all thresholds and the FSS sums of all (threshold, scale) pairs; states are sum-reduced
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


class RainfallMetricSuite:
    def __init__(self, config: Optional[MetricConfig] = None):
        cfg = config or MetricConfig()
        if len(cfg.thresholds) > 4 or len(cfg.scales) > 4 or any(int(s) not in (1, 2, 4, 8) for s in cfg.scales):
            raise ValueError("RainfallMetricSuite supports up to 4 thresholds and up to 4 scales out of {1,2,4,8}")
        self.cfg = cfg
        self.thr = [float(t) for t in cfg.thresholds]
        self.scales = [int(s) for s in cfg.scales]
        self._thr_c = (ctypes.c_float * len(self.thr))(*self.thr)
        self._sc_c = (ctypes.c_int * len(self.scales))(*self.scales)
        self.device: Optional[torch.device] = None
        self.state = None
        self.scratch = None

    def to(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RainfallMetricSuite runs on CUDA (sm_100a) only; there is no CPU path")
        self.state = torch.zeros(51, dtype=torch.float32, device=self.device)
        self.scratch = torch.zeros(50, dtype=torch.float64, device=self.device)
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
        LIB.call("p2i_metrics_update", ptr(p), ptr(t), N, H, W, self._thr_c, len(self.thr), self._sc_c, len(self.scales),
                 1 if self.cfg.apply_transform else 0, ptr(self.scratch), ptr(self.state), stream())
        if H > 10 and W > 10:
            for i in range(0, N, 65535):
                k = min(65535, N - i)
                LIB.call("p2i_ssim_update", ptr(p.view(N, H, W)[i:]), ptr(t.view(N, H, W)[i:]), k, H, W,
                         1 if self.cfg.apply_transform else 0, float(self.cfg.data_range), ptr(self.ssim_state), stream())

    def all_reduce(self, group=None) -> None:
        """Sum the streaming state over data-parallel ranks (the reference declares dist_reduce_fx='sum')."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.state, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.ssim_state, op=dist.ReduceOp.SUM, group=group)

    def compute(self) -> Dict[str, float]:
        s = self.state.detach().cpu()          # host synchronisation: compute() is outside the hot loop
        n = torch.clamp(s[2], min=1.0)
        ss = self.ssim_state.detach().cpu()
        ssim = float(ss[0] / ss[1]) if float(ss[1]) > 0 else float("nan")
        out: Dict[str, float] = {"mae": float(s[0] / n), "rmse": float(torch.sqrt(s[1] / n)), "ssim": ssim}
        for i, thr in enumerate(self.thr):
            h, m, f, c = s[3 + 4 * i], s[4 + 4 * i], s[5 + 4 * i], s[6 + 4 * i]
            pre = f"cat_thr{thr:.2f}"
            out[f"{pre}/pod"] = float(h / (h + m + EPS))
            out[f"{pre}/far"] = float(f / (h + f + EPS))
            out[f"{pre}/csi"] = float(h / (h + m + f + EPS))
            den = (m + f) * (f + c) + (h + m) * (m + c)
            out[f"{pre}/hss"] = float(2 * (h * c - m * f) / (den + EPS))
        for i, thr in enumerate(self.thr):
            for j, sc in enumerate(self.scales):
                cnt = s[35 + 4 * i + j]
                if cnt == 0:
                    continue
                out[f"fss_thr{thr:.2f}_s{sc}"] = float(s[19 + 4 * i + j] / cnt)
        return out

    def reset(self) -> None:
        if self.state is not None:
            self.state.zero_()
            self.scratch.zero_()
            self.ssim_state.zero_()
