"""Yardstick for the LOSS tolerances of the train-step parity tests: how far do the losses of the
reference arithmetic move when only the precision changes?  The oracle's GAN step is run twice on the CPU -- fp32, and
with the generator/discriminator forward under torch bf16 autocast (conv operands AND stored activations rounded to
bf16, fp32 accumulation: the arithmetic our kernels implement) -- on the two test configurations.

    python tests/tools/bf16_loss_yardstick.py            # prints relative deviations per loss term

Results (this container, torch 2.11 CPU) are recorded in tests/test_gpu_training.py and tests/test_gpu_bench_shapes.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import synth  # noqa: E402
from oracle import p2i_oracle as O  # noqa: E402
from p2igan_b200 import build_discriminator, build_generator  # noqa: E402


def losses(H, W, B, n_obs, seed, autocast):
    cfg = synth.make_cfg(H, W)
    torch.manual_seed(2024)
    g_sd = {k: v.detach().clone() for k, v in build_generator(cfg).state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in build_discriminator(cfg).state_dict().items()}
    fr, mf, mk = synth.make_batch(B, 16, H, W, n_obs, seed)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        return O.gan_train_step(g_sd, d_sd, fr, mf, mk, {}, {}, 1, idw="exact")


if __name__ == "__main__":
    for (H, W, B, n_obs, seed) in [(32, 32, 2, 12, 100), (32, 32, 2, 12, 101), (128, 128, 4, 79, 500)]:
        a, b = losses(H, W, B, n_obs, seed, False), losses(H, W, B, n_obs, seed, True)
        print(f"{H}x{W} B={B} seed={seed}: " + "  ".join(
            f"{k} fp32={a[k]:.6g} bf16={b[k]:.6g} rel={abs(a[k] - b[k]) / max(abs(a[k]), 1e-12):.2e}" for k in a))
