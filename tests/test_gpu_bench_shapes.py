"""Parity at the BENCHMARKED shapes: every conv_halo / conv_wgrad instantiation that the
B=16 16x128x128 training step of bench.py launches is compared with torch fp32 convolutions here, with an assertion on WHICH
template ran (p2i_conv_last_variant), and one full GANTrainStep at B=16, 128x128, 79 gauges is compared with the oracle's
step (losses, every parameter gradient, spectral-norm u/v).

Tolerances: conv kernels rel-L2 <= 5e-3 (bf16 operands, fp32 accumulate), wgrad 2e-3; step losses rel 1e-2
(SURVEY.md 8c gate; the bf16-autocast yardstick of the reference arithmetic itself is in tests/tools/bf16_loss_yardstick.py);
per-tensor gradient rel-L2 <= 8e-2 (yardstick tests/tools/bf16_grad_yardstick.py: 6.1e-2); D logits rel-L2 <= 3e-2."""
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import p2i_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B = 16


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _cl(x):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)


def variant():
    from p2igan_b200._lib import LIB
    return int(LIB.load().p2i_conv_last_variant())


def halo(nt, res, cg, mb=1):
    return 2000000 + mb * 100000 + nt * 100 + (10 if res else 0) + cg


# generator level -> (channels, H=W, instantiation the B=16 train step runs for its 3x3 convs)
LEVELS = [
    (64, 128, halo(64, True, 1)),         # resident weights, single CTA
    (128, 64, halo(128, True, 2)),        # resident weights, CTA pairs
    (256, 32, halo(256, False, 2)),       # streamed weights, N = 256 pairs: 31 launches/step, no parity test in round 1
    (512, 16, halo(128, False, 2)),       # too few pairs for N = 256 -> N = 128 pairs
]


@pytest.mark.parametrize("C,HW,expect", LEVELS)
def test_generator_level_convs_at_batch_16(C, HW, expect):
    """ResBlock_do convs of one level at B=16: forward (+ReLU), forward (+residual), data gradient (ReLU mask epilogue) and
    weight gradient, each against F.conv2d / autograd in fp32 on the bf16-rounded operands."""
    from p2igan_b200 import ops
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, HW, HW, generator=g).bfloat16().float()
    r = torch.randn(B, C, HW, HW, generator=g).bfloat16().float()
    dy = torch.randn(B, C, HW, HW, generator=g).bfloat16().float()
    w = (torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5).bfloat16().float()
    w_cl = w.permute(2, 3, 0, 1).reshape(9, C, C).contiguous().to(DEV, torch.bfloat16)
    w_t = w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, C, C).contiguous().to(DEV, torch.bfloat16)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, None, 1, 1)
    y_ref.backward(dy)

    y = ops.conv2d_cl(_cl(x), w_cl, None, True)
    v_relu = variant()
    assert rel_l2(y.float().permute(0, 3, 1, 2), y_ref.detach().relu()) < 5e-3
    y = ops.conv2d_cl(_cl(x), w_cl, _cl(r), False)
    v_res = variant()
    assert rel_l2(y.float().permute(0, 3, 1, 2), y_ref.detach() + r) < 5e-3
    # dgrad with the ReLU mask of the saved activation fused into the epilogue (generator.py eblock_bwd)
    dx = ops.conv2d_cl(_cl(dy), w_t, None, False, mask=_cl(r))
    v_dgrad = variant()
    assert rel_l2(dx.float().permute(0, 3, 1, 2), xr.grad * (r > 0)) < 5e-3
    dW = ops.conv2d_wgrad(_cl(x), _cl(dy), 3)
    v_wgrad = variant()
    torch.cuda.synchronize()
    assert rel_l2(dW, wr.grad.permute(2, 3, 0, 1).reshape(9, C, C)) < 2e-3
    assert (v_relu, v_res, v_dgrad) == (expect, expect, expect), (v_relu, v_res, v_dgrad, expect)
    assert v_wgrad == 3000000 + (128 if C >= 128 else 64) * 100 + (10 if C == 64 else 0), v_wgrad


@pytest.mark.parametrize("Cin,Cout,HW,expect", [
    (512, 256, 16, halo(128, True, 2)), (256, 128, 32, halo(128, True, 2)), (128, 64, 64, halo(64, True, 1)),
    (256, 512, 16, halo(128, True, 2)), (128, 256, 32, halo(256, True, 2)), (64, 128, 64, halo(128, True, 2)),
])
def test_uppos_projections_at_batch_16(Cin, Cout, HW, expect):
    """The 1x1 projections of UPPos (forward: Cin -> Cin/2 at low resolution; data gradient: the transposed weight) and
    their weight gradient at B=16."""
    from p2igan_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, HW, HW, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5).bfloat16().float()
    y = ops.conv2d_cl(_cl(x), w.reshape(1, Cout, Cin).contiguous().to(DEV, torch.bfloat16))
    v = variant()
    assert rel_l2(y.float().permute(0, 3, 1, 2), F.conv2d(x, w)) < 5e-3
    assert v == expect, (v, expect)
    if Cin > Cout:
        dy = torch.randn(B, Cout, HW, HW, generator=g).bfloat16().float()
        dW = ops.conv2d_wgrad(_cl(x), _cl(dy), 1)
        torch.cuda.synchronize()
        ref = torch.einsum("bohw,bihw->oi", dy.double(), x.double()).float()
        assert rel_l2(dW.reshape(Cout, Cin), ref) < 2e-3


class _VariantLog:
    """Records p2i_conv_last_variant() after every tensor-core conv launch of the discriminator / generator."""

    def __init__(self):
        self.seen = {}

    def __enter__(self):
        from p2igan_b200 import disc_bwd, disc_ops, ops
        self.mods = [(disc_ops, "conv_igemm"), (disc_bwd, "conv_igemm"), (ops, "conv2d_cl"), (disc_ops, "conv_wgrad"),
                     (disc_bwd, "conv_wgrad"), (ops, "conv2d_wgrad")]
        self.orig = [getattr(m, n) for m, n in self.mods]
        for (m, n), f in zip(self.mods, self.orig):
            def wrapped(*a, _f=f, **k):
                r = _f(*a, **k)
                v = variant()
                self.seen[v] = self.seen.get(v, 0) + 1
                return r
            setattr(m, n, wrapped)
        return self

    def __exit__(self, *exc):
        for (m, n), f in zip(self.mods, self.orig):
            setattr(m, n, f)


def test_discriminator_forward_backward_at_batch_16():
    """P2IDiscriminator at the bench shape (B=16, 16x128x128): logits, every parameter gradient (spectral-norm backward
    included) and the input gradient vs the oracle's autograd; the launches must include the N=256 pair kernels."""
    from p2igan_b200 import build_discriminator
    from p2igan_b200.losses import gan_loss
    torch.manual_seed(2024)
    D = build_discriminator(synth.make_cfg(128, 128))
    gp = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in D.named_parameters():
            if n.endswith("bias") or n.startswith("alpha"):
                p.add_(torch.randn(p.shape, generator=gp) * 0.1)
    sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    frames = synth.make_batch(B, 16, 128, 128, 79, 21)[0]
    x_ref = frames.clone().requires_grad_(True)
    train = [k for k in sd if k.endswith("weight_orig") or k.endswith("bias") or k == "alpha2d"]
    p = {k: (v.clone().requires_grad_(True) if k in train else v.clone()) for k, v in sd.items()}
    out_ref = O.discriminator_forward(p, x_ref, training=True)
    loss_ref = O.gan_loss(out_ref, True, "hinge", True) + 0.3 * (out_ref ** 2).mean()
    g_all = torch.autograd.grad(loss_ref, [p[k] for k in train] + [x_ref])
    g_ref, gx_ref = dict(zip(train, g_all[:-1])), g_all[-1]

    D = D.to(DEV).train()
    x = frames.to(DEV).requires_grad_(True)
    with _VariantLog() as log:
        out = D(x)
        loss = gan_loss(out, True, loss_type="hinge", is_disc=True) + 0.3 * (out ** 2).mean()
        loss.backward()
    torch.cuda.synchronize()
    report = {"logits": rel_l2(out.detach(), out_ref.detach()), "loss": abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)),
              "input": rel_l2(x.grad, gx_ref)}
    for n, prm in D.named_parameters():
        if n != "alpha3d":
            report[n] = rel_l2(prm.grad, g_ref[n])
    bad = {k: v for k, v in report.items() if v > (3e-2 if k == "logits" else (1e-2 if k == "loss" else 8e-2))}
    assert not bad, (bad, report)
    assert halo(256, False, 2) in log.seen and halo(64, True, 1) in log.seen, log.seen
    assert all(v // 1000000 in (2, 3) for v in log.seen), log.seen          # halo + first-generation wgrad kernels only


def test_gan_train_step_at_bench_config_matches_oracle():
    """ONE full G+D iteration at BASELINE configs[2] (B=16, 16x128x128, 79 gauges, hinge + weighted-L1 + temporal KL, Adam)
    in scripts/train.py's order (:240-326) vs the oracle step with the kernel's IDW tie rule: losses, per-tensor
    gradients of G and D, u/v after the step's three power iterations, and the kernel variants that ran."""
    from p2igan_b200 import build_discriminator, build_generator
    from p2igan_b200.train_step import GANTrainStep
    cfg = synth.make_cfg(128, 128)
    torch.manual_seed(2024)
    G, D = build_generator(cfg), build_discriminator(cfg)
    g_sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    fr, mf, mk = synth.make_batch(B, 16, 128, 128, 79, 1000)
    G, D = G.to(DEV).train(), D.to(DEV).train()
    ts = GANTrainStep(cfg, G, D)
    with _VariantLog() as log:
        ours = {k: float(v) for k, v in ts.step(fr.to(DEV), mf.to(DEV), mk.to(DEV)).items()}
    torch.cuda.synchronize()
    grads = {}
    ref = O.gan_train_step(g_sd, d_sd, fr, mf, mk, {}, {}, 1, idw="exact", grads_out=grads)
    report = {}
    for k in ("rec", "pool", "reg", "dis", "total"):
        report["loss." + k] = abs(ours[k] - ref[k]) / max(abs(ref[k]), 1e-12)
    report["loss.adv_abs"] = abs(ours["adv"] - ref["adv"])
    bad = {k: v for k, v in report.items() if v > (2e-3 if k == "loss.adv_abs" else 1e-2)}
    # gradients: the flat buffers still hold this step's (un-averaged, world size 1) gradients after the Adam updates
    for n, prm in G.named_parameters():
        if prm.requires_grad:
            report["G." + n] = r = rel_l2(prm.grad, grads["g"][n])
            if r > 8e-2:
                bad["G." + n] = r
    for n, prm in D.named_parameters():
        if n != "alpha3d":
            report["D." + n] = r = rel_l2(prm.grad, grads["d"][n])
            if r > 8e-2:
                bad["D." + n] = r
    got_d = D.state_dict()
    for k in d_sd:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            e = float((got_d[k].cpu() - d_sd[k]).abs().max())
            if e > 1e-3:
                bad["uv." + k] = e
    worst = sorted(((v, k) for k, v in report.items()), reverse=True)[:6]
    print("bench-config train step: losses", {k: (ours[k], ref[k]) for k in ours}, "worst deviations", worst)
    assert not bad, (bad, worst)
    for need in (halo(256, False, 2), halo(256, True, 2), halo(128, False, 2), halo(128, True, 2), halo(64, True, 1),
                 halo(64, True, 2)):
        assert need in log.seen, (need, log.seen)
    assert all(v // 1000000 in (2, 3) for v in log.seen), log.seen
    # parameters moved by Adam exactly where the oracle's moved (beta1 = 0: lr-sized steps; bounded by 2.5 lr per element)
    got_g = G.state_dict()
    for k, v in g_sd.items():
        assert float((got_g[k].cpu() - v).abs().max()) <= 2.5e-4, k
