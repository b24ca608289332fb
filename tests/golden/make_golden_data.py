"""Golden vectors for the batch-preparation row by EXECUTING THE REFERENCE's host code
(p2igan_bench/data/sti_dataset.py: create_mask 'stis' + Dataset.post_process + _crop_center) -- build container only.

    python tests/golden/make_golden_data.py       # writes tests/golden/reference_batch_prep.pt

decord / h5py / zarr are absent from this image; they are imported at module level by the reference but not used by
the functions exercised here, so empty stub modules are installed first.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
for name in ("decord", "h5py", "zarr"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.VideoReader = object
        sys.modules[name] = m

from p2igan_bench.data import sti_dataset as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.RandomState(11)
    T, H0, W0, H, W = 6, 20, 24, 16, 16
    video = rng.randint(0, 256, size=(T, H0, W0), dtype=np.uint8)
    mask2d = (rng.rand(H0, W0) < 0.1).astype(np.float64)
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        np.savetxt(f, mask2d, fmt="%d")
        mask_file = f.name
    ds = S.Dataset.__new__(S.Dataset)            # post_process only reads these attributes
    ds.sample_length, ds.transform = 4, S.transform
    ds.mask_type, ds.mask_file, ds.block_sizes, ds.mask_keep, ds.mask_interval = "stis", mask_file, [4], 4, [2, 5]
    ds.height, ds.width = H, W
    frames, masked, mask = ds.post_process(video[..., np.newaxis])
    os.unlink(mask_file)
    out = {"video_u8": torch.from_numpy(video), "mask2d_u8": torch.from_numpy(mask2d.astype(np.uint8)), "sample_length": 4,
           "H": H, "W": W, "frames": frames.float(), "masked": masked.float(), "mask": mask.float()}
    print({k: (tuple(v.shape), str(v.dtype)) if torch.is_tensor(v) else v for k, v in out.items()})
    torch.save(out, os.path.join(HERE, "reference_batch_prep.pt"))


if __name__ == "__main__":
    main()
