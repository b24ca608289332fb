"""Generate golden vectors by EXECUTING THE REFERENCE (CPU, fp32) -- run in the build container only.

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

/root/reference is imported unmodified; the packages it needs but this image lacks are replaced by
in-memory stubs (torchmetrics.Metric / StructuralSimilarityIndexMeasure), exactly as SURVEY.md 8c
describes.  Nothing here is imported by the product or by the GPU tests: the tests read only the
committed ``*.pt`` files.  Inputs are produced by ``tests/synth.py`` (seeded), so the tests can
regenerate them without the reference.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")


def _install_stubs():
    tm = types.ModuleType("torchmetrics")

    class Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default.clone()
            self.register_buffer(name, default.clone())

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

    class SSIM(Metric):
        def __init__(self, data_range=1.0):
            super().__init__()

        def update(self, a, b):
            pass

        def compute(self):
            return torch.tensor(float("nan"))

    tm.Metric = Metric
    img = types.ModuleType("torchmetrics.image")
    img.StructuralSimilarityIndexMeasure = SSIM
    tm.image = img
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.image"] = img


_install_stubs()

from p2igan_bench.models import build_discriminator, build_generator  # noqa: E402  (reference)
from p2igan_bench.modules import ReconstructionLoss, gan_loss  # noqa: E402
from p2igan_bench.modules.layer import InputBlock  # noqa: E402
from p2igan_bench.metrics.metric import MetricConfig, RainfallMetricSuite  # noqa: E402

import synth  # noqa: E402  (tests/synth.py)


def sd_fingerprint(sd):
    out = {}
    for k, v in sd.items():
        f = v.detach().double().reshape(-1)
        out[k] = (tuple(v.shape), float(f.sum()), float(f.abs().sum()), [float(x) for x in f[:3]])
    return out


def sub(t, n=4096):
    """Deterministic strided subsample of a tensor (flattened) to keep fixtures small."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].clone()


def run_generator(cfg, B, seed_model, seed_data, n_obs, tie_free, keep_full):
    torch.manual_seed(seed_model)
    G = build_generator(cfg)
    H, W = cfg["data"]["train"]["h"], cfg["data"]["train"]["w"]
    frames, masked, masks = synth.make_batch(B, 16, H, W, n_obs, seed_data, tie_free=tie_free)
    inter = {}

    def hook(name):
        def fn(m, i, o):
            inter[name] = o.detach().clone()
        return fn

    hs = [G.input.register_forward_hook(hook("input")), G.Convsin[0].register_forward_hook(hook("convsin")),
          G.ConvsOut[0].register_forward_hook(hook("z"))]
    for l in range(4):
        hs.append(G.Decoder[l].register_forward_hook(hook(f"dec{l}")))
    for l in range(3):
        hs.append(G.UP[l].register_forward_hook(hook(f"up{l}")))
    G.eval()
    with torch.no_grad():
        out = G(masked, masks)
    for h in hs:
        h.remove()
    rec = {"out": out if keep_full else sub(out), "fingerprint": sd_fingerprint(G.state_dict())}
    for k, v in inter.items():
        rec[k] = v if (keep_full and v.numel() <= 70000) else sub(v)
    rec["full"] = keep_full
    return G, rec, (frames, masked, masks)


def main():
    torch.set_num_threads(8)
    cfg32 = synth.make_cfg(32, 32)
    cfg128 = synth.make_cfg(128, 128)
    gold = {"torch": torch.__version__}

    # ---- generator, 32x32 (full tensors) : standard repeated mask (ties) and jittered (tie-free)
    G, rec, (frames, masked, masks) = run_generator(cfg32, 2, 2024, 1, 12, False, True)
    gold["g32"] = rec
    _, rec_tf, _ = run_generator(cfg32, 2, 2024, 2, 12, True, True)
    gold["g32_tiefree"] = rec_tf

    # ---- generator trunk on a *given* InputBlock output, perturbed weights (D != 0, pos != 0)
    torch.manual_seed(7)
    G2 = build_generator(cfg32)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in G2.named_parameters():
            if n.endswith(".D") or n.endswith(".pos") or n.endswith("proj.bias") or n.endswith("conv.bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    G2.eval()
    with torch.no_grad():
        o2 = G2(masked, masks)
    gold["g32_perturbed"] = {"out": o2, "seed_model": 7, "seed_perturb": 11}

    # ---- generator, 128x128, B=1 (subsampled)
    _, rec128, (f128, m128, k128) = run_generator(cfg128, 1, 2024, 1, 79, False, False)
    gold["g128"] = rec128

    # ---- IDW alone at 128 (reference InputBlock with identity gates is not needed: use module as is)
    torch.manual_seed(3)
    ib = InputBlock(depth=2, k=4, rho=2.0, tau=0.05, chunk=16384)
    with torch.no_grad():
        ib_out = ib(m128.reshape(1, 16, 128, 128), k128.reshape(1, 16, 128, 128))
    gold["inputblock128"] = {"out": ib_out[0, ::5, ::4, ::4].clone(), "seed_model": 3,
                             "sd": {k: v.clone() for k, v in ib.state_dict().items()}}

    # ---- discriminator (train mode: power iteration mutates u/v) and eval mode
    torch.manual_seed(2024)
    D = build_discriminator(cfg32)
    fp0 = sd_fingerprint(D.state_dict())
    D.train()
    with torch.no_grad():
        l_train = D(frames)
    sd_after = {k: v.clone() for k, v in D.state_dict().items() if k.endswith("_u") or k.endswith("_v")}
    D.eval()
    with torch.no_grad():
        l_eval = D(frames)
    gold["d32"] = {"fingerprint": fp0, "logits_train": l_train, "logits_eval": l_eval,
                   "uv_after": {k: sub(v, 64) for k, v in sd_after.items()}}
    torch.manual_seed(2024)
    D128 = build_discriminator(cfg128)
    D128.train()
    with torch.no_grad():
        l128 = D128(f128)
    gold["d128"] = {"logits_train": l128}

    # ---- losses
    pred = gold["g32"]["out"]
    rl = ReconstructionLoss(k1_alpha=0.05)
    loss, parts = rl(pred, frames, masks)
    gold["loss32"] = {"total": float(loss), "pool": parts["pool"], "reg": parts["reg"],
                      "hinge_d_real": float(gan_loss(l_train, True, loss_type="hinge", is_disc=True)),
                      "hinge_d_fake": float(gan_loss(l_train, False, loss_type="hinge", is_disc=True)),
                      "hinge_g": float(gan_loss(l_train, True, loss_type="hinge", is_disc=False)),
                      "lsgan_real": float(gan_loss(l_train, True, loss_type="lsgan")),
                      "nsgan_fake": float(gan_loss(torch.sigmoid(l_train), False, loss_type="nsgan"))}

    # ---- metrics (dBZ-like scale so thresholds bite), two updates
    suite = RainfallMetricSuite(MetricConfig())
    suite.update(pred * 85.0, frames * 85.0)
    suite.update(frames.flip(1) * 85.0, frames * 85.0)
    m = suite.compute()
    m.pop("ssim", None)
    gold["metrics32"] = m

    # ---- one full GAN training step in train.py's order (B=2, 32x32), then a second one
    torch.manual_seed(2024)
    G = build_generator(cfg32)
    D = build_discriminator(cfg32)
    og = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.0, 0.99))
    od = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.99))
    G.train(); D.train()
    steps = []
    for it in range(2):
        fr, mf, mk = synth.make_batch(2, 16, 32, 32, 12, 100 + it)
        preds = G(mf, mk)
        loss_g, parts = rl(preds, fr, mk)
        for p in D.parameters():
            p.requires_grad_(True)
        lf = D(preds.detach()); lr_ = D(fr)
        loss_d = (gan_loss(lr_, True, loss_type="hinge", is_disc=True)
                  + gan_loss(lf, False, loss_type="hinge", is_disc=True)) * 0.5
        od.zero_grad(); loss_d.backward()
        d_gn = {n: float(p.grad.double().norm()) for n, p in D.named_parameters() if p.grad is not None}
        od.step()
        for p in D.parameters():
            p.requires_grad_(False)
        adv = gan_loss(D(preds), True, loss_type="hinge", is_disc=False) * 0.01
        total = loss_g + adv
        og.zero_grad(); total.backward()
        g_gn = {n: float(p.grad.double().norm()) for n, p in G.named_parameters() if p.grad is not None}
        og.step()
        for p in D.parameters():
            p.requires_grad_(True)
        steps.append({"rec": float(loss_g), "pool": parts["pool"], "reg": parts["reg"], "adv": float(adv),
                      "dis": float(loss_d), "g_grad_norms": g_gn, "d_grad_norms": d_gn})
    gold["train32"] = {"steps": steps, "g_after": sd_fingerprint(G.state_dict()),
                       "d_after": sd_fingerprint(D.state_dict())}

    path = os.path.join(HERE, "reference_golden.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
