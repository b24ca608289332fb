"""Golden values of the paper-evaluation metrics by EXECUTING THE REFERENCE's experiments/exp1.py
(run_exp1: transform_mmhr, centre crop, gauge / radar pixel selection, MAE, RMSE, PSS, pooled SSIM, delta-TSSIM, NSE, and
the categorical scores) on seeded arrays -- build container only.

    python tests/golden/make_golden_exp1.py      # writes tests/golden/reference_exp1.pt
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from experiments import exp1  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.RandomState(21)
    T, H, W, crop = 12, 72, 80, 64
    truth = (rng.rand(T, H, W) ** 3 * 255.0).astype(np.float32)            # 0..255 "dBZ-like" counts, as the saved zarrs hold
    pred = np.clip(truth + rng.randn(T, H, W).astype(np.float32) * 12.0, 0, 255).astype(np.float32)
    mask = rng.rand(crop, crop) < 0.03
    out = {"truth": torch.from_numpy(truth), "pred": torch.from_numpy(pred), "mask": torch.from_numpy(mask), "crop": crop, "results": {}}
    for mode in ("radar", "gauge"):
        for div3 in (True, False):
            r = exp1.run_exp1({"m": pred}, truth, mask, mode, crop, use_pool8=True, divide_by_3=div3)["m"]
            out["results"][f"{mode}_{int(div3)}"] = r
    # all pixels (empty gauge mask, "radar" = the complement), no /3, full frame cropped to the square the metric suite sees:
    # the categorical scores are then directly comparable with RainfallMetricSuite (metrics/metric.py) on the same arrays
    none = np.zeros((crop, crop), dtype=bool)
    out["results"]["all_0"] = exp1.run_exp1({"m": pred}, truth, none, "radar", crop, use_pool8=True, divide_by_3=False)["m"]
    print(out["results"]["radar_1"])
    torch.save(out, os.path.join(HERE, "reference_exp1.pt"))


if __name__ == "__main__":
    main()
