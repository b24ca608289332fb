"""CPU: oracle/paper_metrics.py vs the golden produced by the reference's own experiments/exp1.run_exp1 (row N4)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import paper_metrics as PM  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "reference_exp1.pt")


def _close(a, b, tol):
    return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= tol * max(1.0, abs(b))


def test_paper_metrics_match_reference_exp1():
    g = torch.load(GOLD, weights_only=False)
    pred, truth, mask, crop = g["pred"].numpy(), g["truth"].numpy(), g["mask"].numpy(), g["crop"]
    import numpy as np
    cases = {"radar_1": ("radar", True, mask), "radar_0": ("radar", False, mask), "gauge_1": ("gauge", True, mask),
             "gauge_0": ("gauge", False, mask), "all_0": ("radar", False, np.zeros_like(mask))}
    for key, (mode, div3, m) in cases.items():
        ref = g["results"][key]
        ours = PM.run_exp1(pred, truth, m, mode, crop, divide_by_3=div3)
        for k, v in ref.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    assert _close(ours[k][kk], vv, 1e-9), (key, k, kk, ours[k][kk], vv)
            else:
                tol = 2e-5 if k in ("SSIM", "DTSSIM_L1", "DTSSIM_L2") else 1e-9          # float32 means in torch vs numpy
                assert _close(ours[k], v, tol), (key, k, ours[k], v)
