"""Second opinion on SSIM.  torchmetrics 1.0.3 -- what the reference's RegressionMetrics calls
(p2igan_bench/metrics/metric.py:36,55-56,69) -- is absent from this image, so ``oracle.ssim_per_image`` (the restatement
the CUDA kernel is tested against) stays labelled "parity unpinned against torchmetrics".  This test pins it against an
INDEPENDENT implementation written from the SSIM paper (Wang et al. 2004: 11x11 gaussian window, sigma 1.5, K1 0.01,
K2 0.03, biased local moments) with scipy.ndimage separable correlations in float64 -- no code shared with the oracle or
the kernel -- averaged over the interior where the window fits (the region torchmetrics keeps after cropping its
reflect-padded maps)."""
import numpy as np
import pytest
import torch

from oracle import p2i_oracle as O

ndimage = pytest.importorskip("scipy.ndimage")


def ssim_paper(x: np.ndarray, y: np.ndarray, data_range: float = 1.0, size: int = 11, sigma: float = 1.5) -> float:
    x, y = x.astype(np.float64), y.astype(np.float64)
    t = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-(t * t) / (2.0 * sigma * sigma))
    g /= g.sum()

    def blur(a):
        return ndimage.correlate1d(ndimage.correlate1d(a, g, axis=0, mode="constant"), g, axis=1, mode="constant")

    mx, my = blur(x), blur(y)
    sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    m = ((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sxx + syy + c2))
    h = size // 2
    return float(m[h:-h, h:-h].mean())         # interior: the window lies inside the image, padding never enters


@pytest.mark.parametrize("shape,scale,data_range", [((3, 37, 53), 1.0, 1.0), ((2, 128, 128), 1.0, 1.0), ((2, 64, 48), 12.0, 1.0),
                                                    ((1, 40, 40), 255.0, 255.0)])
def test_oracle_ssim_matches_independent_paper_implementation(shape, scale, data_range):
    g = torch.Generator().manual_seed(shape[1] + int(scale))
    pred = torch.rand(shape, generator=g) ** 2 * scale
    tgt = (pred + 0.25 * scale * torch.rand(shape, generator=g)).clamp(0, scale)
    ours = O.ssim_per_image(pred[:, None], tgt[:, None], data_range=data_range)
    for i in range(shape[0]):
        ref = ssim_paper(pred[i].numpy(), tgt[i].numpy(), data_range)
        assert abs(float(ours[i]) - ref) < 2e-5, (i, float(ours[i]), ref)


def test_ssim_limits():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 1, 32, 32, generator=g)
    assert abs(float(O.ssim_per_image(x, x)) - 1.0) < 1e-6                     # identical images
    assert float(O.ssim_per_image(x, 1.0 - x)) < 0.1                            # anti-correlated images
