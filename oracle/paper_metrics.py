"""CPU restatement of the paper-evaluation metrics of the reference's experiments/exp1.py -- TEST INFRASTRUCTURE ONLY
("cross-checks").  numpy + a little torch, each function citing the lines it follows; pinned by
tests/golden/reference_exp1.pt, produced by running the reference's own run_exp1 (tests/golden/make_golden_exp1.py).
Imported only by tests/."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def transform_mmhr(arr, divide_by_3: bool = True):
    """exp1.py:147-155: floor at 0.001, optional /3, R = 0.036 * 10^(x/16) with the exponent capped at 38, clip to [0, 200]."""
    a = np.maximum(np.asarray(arr, dtype=np.float64), 0.001)
    if divide_by_3:
        a = a / 3.0
    return np.clip(10 ** np.clip(a * 0.0625, None, 38.0) * 0.036, 0.0, 200.0)


def crop_center(a, size: int):
    """experiments/io.py:23-30."""
    t, h, w = a.shape
    top, left = (h - size) // 2, (w - size) // 2
    return a[:, top:top + size, left:left + size]


def select(a, mask, invert: bool):
    """experiments/io.py:115-123."""
    m = ~mask.astype(bool) if invert else mask.astype(bool)
    return a.reshape(a.shape[0], -1)[:, m.ravel()]


def pss(pred, gt, bins: int = 50, min_value: float = 0.5) -> float:
    """exp1.py:20-62: mean over frames of the histogram overlap of the values above min_value (shared range)."""
    pred, gt = np.asarray(pred, np.float32), np.asarray(gt, np.float32)
    both = np.concatenate([pred.ravel(), gt.ravel()])
    both = both[np.isfinite(both)]
    both = both[both > min_value]
    if both.size == 0:
        return float("nan")
    vmin, vmax = float(both.min()), float(both.max())
    if vmin == vmax:
        vmax = vmin + 1e-6
    scores = []
    for p, g in zip(pred, gt):
        p, g = p.ravel(), g.ravel()
        p, g = p[np.isfinite(p) & (p > min_value)], g[np.isfinite(g) & (g > min_value)]
        if p.size == 0 or g.size == 0:
            continue
        ph, _ = np.histogram(p, bins=bins, range=(vmin, vmax))
        gh, _ = np.histogram(g, bins=bins, range=(vmin, vmax))
        scores.append(float(np.minimum(ph / (ph.sum() + 1e-12), gh / (gh.sum() + 1e-12)).sum()))
    return float(np.mean(scores)) if scores else float("nan")


def _pool8(x):
    """exp1.py:85-88 (avg_pool2d k=8 s=8 on float32)."""
    t, h, w = x.shape
    h8, w8 = h // 8, w // 8
    return x[:, :h8 * 8, :w8 * 8].astype(np.float32).reshape(t, h8, 8, w8, 8).mean(axis=(2, 4), dtype=np.float32)


def _ssim_global(a, b, c1=0.01 ** 2, c2=0.03 ** 2):
    """exp1.py:91-99: single-window SSIM of two fields (float32 arithmetic, as torch)."""
    a, b = a.astype(np.float32), b.astype(np.float32)
    mu_a, mu_b = a.mean(dtype=np.float32), b.mean(dtype=np.float32)
    sa, sb = ((a - mu_a) ** 2).mean(dtype=np.float32), ((b - mu_b) ** 2).mean(dtype=np.float32)
    sab = ((a - mu_a) * (b - mu_b)).mean(dtype=np.float32)
    return (2 * mu_a * mu_b + c1) * (2 * sab + c2) / ((mu_a ** 2 + mu_b ** 2 + c1) * (sa + sb + c2) + 1e-10)


def ssim_spatial(pred, gt, use_pool8: bool = True) -> float:
    """exp1.py:110-121."""
    if use_pool8:
        pred, gt = _pool8(pred), _pool8(gt)
    return float(np.mean([_ssim_global(p, g) for p, g in zip(pred, gt)], dtype=np.float32))


def delta_tssim(pred, gt, lag: int = 1, use_pool8: bool = True) -> float:
    """exp1.py:102-107,124-135: mean over t of SSIM(x_t, x_{t-lag}) of the prediction minus that of the truth."""
    if pred.shape[0] <= lag:
        return float("nan")
    if use_pool8:
        pred, gt = _pool8(pred), _pool8(gt)
    sp = np.array([_ssim_global(pred[t], pred[t - lag]) for t in range(lag, pred.shape[0])], dtype=np.float32)
    sg = np.array([_ssim_global(gt[t], gt[t - lag]) for t in range(lag, gt.shape[0])], dtype=np.float32)
    return float((sp - sg).mean(dtype=np.float32))


def nse(pred, gt) -> float:
    """exp1.py:138-141."""
    return float(1.0 - np.sum((pred - gt) ** 2) / (np.sum((gt - np.mean(gt)) ** 2) + 1e-10))


def categorical(pred, gt, thr: float) -> Dict[str, float]:
    """exp1.py:158-175 (HSS with exp1's own denominator, which differs from metric.py's)."""
    pb, gb = pred >= thr, gt >= thr
    h, m = float((pb & gb).sum()), float((~pb & gb).sum())
    f, c = float((pb & ~gb).sum()), float((~pb & ~gb).sum())
    hss = 2 * (h * c - m * f) / (m ** 2 + f ** 2 + 2 * h * c + (m + f) * (h + c) + 1e-10) if (h + m + f + c) > 0 else float("nan")
    return {"POD": h / (h + m + 1e-10), "FAR": f / (h + f + 1e-10), "CSI": h / (h + m + f + 1e-10), "HSS": hss}


def run_exp1(pred, truth, mask, mode: str, crop: int, thresholds: Tuple[float, ...] = (0.5, 2.0, 4.0, 8.0), use_pool8: bool = True,
             divide_by_3: bool = True) -> Dict:
    """exp1.py:191-242 for one method and array inputs [T,H,W]."""
    truth = crop_center(transform_mmhr(truth, divide_by_3), crop)
    pred = crop_center(transform_mmhr(pred, divide_by_3), crop)
    inv = mode == "radar"
    ps, gs = select(pred, mask, inv), select(truth, mask, inv)
    out = {"MAE": float(np.mean(np.abs(ps - gs))), "RMSE": float(np.sqrt(np.mean((ps - gs) ** 2))), "PSS": pss(ps, gs),
           "SSIM": ssim_spatial(pred, truth, use_pool8), "DTSSIM_L1": delta_tssim(pred, truth, 1, use_pool8),
           "DTSSIM_L2": delta_tssim(pred, truth, 2, use_pool8), "NSE": nse(ps, gs)}
    for thr in thresholds:
        out[f"CAT_{thr:g}"] = categorical(ps, gs, thr)
    return out
