"""CPU: the oracle's batch preparation vs the golden produced by the reference's own host code
(tests/golden/make_golden_data.py: create_mask('stis') + Dataset.post_process + _crop_center)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import p2i_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "reference_batch_prep.pt")


def test_oracle_batch_post_process_is_bit_exact_with_reference():
    g = torch.load(GOLD, weights_only=True)
    T = g["sample_length"]
    video = g["video_u8"][:T].numpy()[None]                      # [1, T, H0, W0] (sample_length truncation, :205-207)
    fr, mf, mk = O.batch_post_process(video, g["mask2d_u8"].numpy(), g["H"], g["W"])
    # reference layout per sample is [T, H, W, 1]; _prepare_batch permutes to [T, 1, H, W]
    for ours, ref in ((fr, g["frames"]), (mf, g["masked"]), (mk, g["mask"])):
        assert np.array_equal(ours[0], ref.permute(0, 3, 1, 2).numpy())
