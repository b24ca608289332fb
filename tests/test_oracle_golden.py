"""CPU: the oracle (oracle/p2i_oracle.py) against vectors produced by EXECUTING the reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare CUDA vs oracle."""
import torch

import synth
from oracle import p2i_oracle as O


def _state(seed, h, w):
    """Reference-identical random init comes from OUR module constructors (checked in test_state_dict)."""
    from p2igan_b200 import build_generator
    torch.manual_seed(seed)
    return {k: v.detach().clone() for k, v in build_generator(synth.make_cfg(h, w)).state_dict().items()}


def _sub(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n]


def _cmp(a, b, full):
    return (a, b) if full and a.shape == b.shape else (_sub(a), b)


def test_generator_32_reference_style_idw_matches_reference(golden):
    g = golden["g32"]
    sd = _state(2024, 32, 32)
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 1)
    out, inter = O.generator_forward(sd, masked, masks, idw="ref", return_intermediates=True)
    # same library calls in the same order => essentially bit-exact
    assert torch.allclose(inter["input"], g["input"], atol=1e-6)
    assert torch.allclose(out, g["out"], atol=2e-5)
    for k in ("dec3", "up2", "dec2", "dec1", "dec0", "z"):
        a, b = _cmp(inter[k], g[k], True)
        assert torch.allclose(a, b, atol=1e-4, rtol=1e-4), k


def test_generator_32_tiefree_exact_idw_close_to_reference(golden):
    g = golden["g32_tiefree"]
    sd = _state(2024, 32, 32)
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 2, tie_free=True)
    x_exact = O.input_block(sd, masked.reshape(2, 16, 32, 32), masks.reshape(2, 16, 32, 32), idw="exact")
    mk = masks.reshape(2, 16, 32, 32)
    for b in range(2):
        tz, ty, tx = O.observed_points(mk[b])
        tie = O.idw_tie_mask(tz, ty, tx, (16, 32, 32))
        # non-tie queries: the reference's cdist-matmul noise bounds the difference (3.9e-3 max)
        d = (x_exact[b] - g["input"][b]).abs()
        assert float(d[~tie].max()) < 5e-3
        assert float(tie.float().mean()) < 0.12


def test_idw_tie_queries_pick_a_valid_subset(golden):
    g = golden["g32"]
    sd = _state(2024, 32, 32)
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 1)
    mf, mk = masked.reshape(2, 16, 32, 32), masks.reshape(2, 16, 32, 32)
    proc = O.gated_frames(sd, mf)
    tz, ty, tx = O.observed_points(mk[0])
    vals = proc[0][tz, ty, tx]
    ours = O.idw_exact(tz, ty, tx, vals, (16, 32, 32))
    assert float(O.idw_tie_mask(tz, ty, tx, (16, 32, 32)).float().mean()) > 0.05   # repeated mask => many ties
    assert O.idw_tie_candidates_ok(ours, tz, ty, tx, vals, (16, 32, 32), max_queries=300)
    # ... and so does the reference's own (unspecified-order) answer
    assert O.idw_tie_candidates_ok(g["input"][0], tz, ty, tx, vals, (16, 32, 32), atol=5e-3, max_queries=300)


def test_generator_perturbed_weights(golden):
    g = golden["g32_perturbed"]
    sd = _state(7, 32, 32)
    gen = torch.Generator().manual_seed(11)
    for n in list(sd):
        if n.endswith(".D") or n.endswith(".pos") or n.endswith("proj.bias") or n.endswith("conv.bias"):
            sd[n] = sd[n] + torch.randn(sd[n].shape, generator=gen) * 0.05
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 1)
    out = O.generator_forward(sd, masked, masks, idw="ref")
    assert torch.allclose(out, g["out"], atol=5e-5)


def test_generator_128_subsampled(golden):
    g = golden["g128"]
    sd = _state(2024, 128, 128)
    frames, masked, masks = synth.make_batch(1, 16, 128, 128, 79, 1)
    out, inter = O.generator_forward(sd, masked, masks, idw="ref", return_intermediates=True)
    assert torch.allclose(_sub(out), g["out"], atol=5e-5)
    assert torch.allclose(_sub(inter["z"]), g["z"], atol=2e-4, rtol=1e-4)


def test_inputblock_128(golden):
    g = golden["inputblock128"]
    sd = {"input.layers." + k.split("layers.")[1]: v for k, v in g["sd"].items()}
    frames, masked, masks = synth.make_batch(1, 16, 128, 128, 79, 1)
    mf, mk = masked.reshape(1, 16, 128, 128), masks.reshape(1, 16, 128, 128)
    ref_style = O.input_block(sd, mf, mk, idw="ref")
    assert torch.allclose(ref_style[0, ::5, ::4, ::4], g["out"], atol=1e-6)
    exact = O.input_block(sd, mf, mk, idw="exact")
    tz, ty, tx = O.observed_points(mk[0])
    tie = O.idw_tie_mask(tz, ty, tx, (16, 128, 128))[::5, ::4, ::4]
    d = (exact[0, ::5, ::4, ::4] - g["out"]).abs()
    assert float(d[~tie].max()) < 5e-3


def _d_state(seed, h, w):
    from p2igan_b200 import build_discriminator
    torch.manual_seed(seed)
    return {k: v.detach().clone() for k, v in build_discriminator(synth.make_cfg(h, w)).state_dict().items()}


def test_discriminator_32(golden):
    g = golden["d32"]
    sd = _d_state(2024, 32, 32)
    frames, _, _ = synth.make_batch(2, 16, 32, 32, 12, 1)
    lt = O.discriminator_forward(sd, frames, training=True)
    assert torch.allclose(lt, g["logits_train"], atol=1e-5, rtol=1e-4)
    for k, v in g["uv_after"].items():
        assert torch.allclose(_sub(sd[k], 64), v, atol=1e-6), k
    le = O.discriminator_forward(sd, frames, training=False)
    assert torch.allclose(le, g["logits_eval"], atol=1e-5, rtol=1e-4)


def test_discriminator_128(golden):
    sd = _d_state(2024, 128, 128)
    frames, _, _ = synth.make_batch(1, 16, 128, 128, 79, 1)
    lt = O.discriminator_forward(sd, frames, training=True)
    assert torch.allclose(lt, golden["d128"]["logits_train"], atol=1e-5, rtol=1e-4)


def test_losses(golden):
    g = golden["loss32"]
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 1)
    pred = golden["g32"]["out"]
    loss, parts = O.reconstruction_loss(pred, frames, 0.05)
    assert abs(float(loss) - g["total"]) < 1e-6
    assert abs(parts["pool"] - g["pool"]) < 1e-6 and abs(parts["reg"] - g["reg"]) < 1e-6
    lt = golden["d32"]["logits_train"]
    assert abs(float(O.gan_loss(lt, True, "hinge", True)) - g["hinge_d_real"]) < 1e-6
    assert abs(float(O.gan_loss(lt, False, "hinge", True)) - g["hinge_d_fake"]) < 1e-6
    assert abs(float(O.gan_loss(lt, True, "hinge", False)) - g["hinge_g"]) < 1e-6
    assert abs(float(O.gan_loss(lt, True, "lsgan")) - g["lsgan_real"]) < 1e-6
    assert abs(float(O.gan_loss(torch.sigmoid(lt), False, "nsgan")) - g["nsgan_fake"]) < 1e-6


def test_metrics(golden):
    frames, masked, masks = synth.make_batch(2, 16, 32, 32, 12, 1)
    pred = golden["g32"]["out"]
    m = O.MetricSuiteOracle()
    m.update(pred * 85.0, frames * 85.0)
    m.update(frames.flip(1) * 85.0, frames * 85.0)
    ours = m.compute()
    ref = golden["metrics32"]
    assert set(ours) == set(ref)
    for k, v in ref.items():
        assert abs(ours[k] - v) < 1e-5 * max(1.0, abs(v)), (k, ours[k], v)


def test_train_step(golden):
    g = golden["train32"]
    from p2igan_b200 import build_discriminator, build_generator
    torch.manual_seed(2024)
    cfg = synth.make_cfg(32, 32)
    g_sd = {k: v.detach().clone() for k, v in build_generator(cfg).state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in build_discriminator(cfg).state_dict().items()}
    og, od = {}, {}
    for it in range(2):
        fr, mf, mk = synth.make_batch(2, 16, 32, 32, 12, 100 + it)
        r = O.gan_train_step(g_sd, d_sd, fr, mf, mk, og, od, it + 1, idw="ref")
        ref = g["steps"][it]
        for k in ("rec", "pool", "reg", "adv", "dis"):
            assert abs(r[k] - ref[k]) < 2e-5 * max(1.0, abs(ref[k])), (it, k, r[k], ref[k])
    for k, (shape, s, a, head) in g["g_after"].items():
        assert abs(float(g_sd[k].double().abs().sum()) - a) < 1e-4 * max(1.0, a), k
    for k, (shape, s, a, head) in g["d_after"].items():
        assert abs(float(d_sd[k].double().abs().sum()) - a) < 1e-4 * max(1.0, a), k
