"""CPU: constructors reproduce the reference's state_dict layout AND its random-init values
(same parameter registration order + same init calls => same RNG stream), checked against
fingerprints recorded from the reference modules (tests/golden/make_golden.py)."""
import torch

import synth


def _check(sd, fp):
    assert list(sd.keys()) == list(fp.keys())
    for k, (shape, s, a, head) in fp.items():
        v = sd[k].detach().double().reshape(-1)
        assert tuple(sd[k].shape) == tuple(shape), k
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k
        assert [float(x) for x in v[:3]] == head, k


def test_generator_state_dict_matches_reference(golden):
    from p2igan_b200 import build_generator
    torch.manual_seed(2024)
    G = build_generator(synth.make_cfg(32, 32))
    sd = G.state_dict()
    assert len(sd) == 113
    _check(sd, golden["g32"]["fingerprint"])
    n_train = sum(p.numel() for p in G.parameters() if p.requires_grad)
    n_all = sum(p.numel() for p in G.parameters())
    # 128x128 has 25 887 984 trainable; only the three `pos` maps depend on H, W
    assert n_train == 25887984 - (128 * 128 + 64 * 64 + 32 * 32) + (32 * 32 + 16 * 16 + 8 * 8)
    assert n_all - n_train == 623376      # frozen D_diag


def test_generator_128_fingerprint(golden):
    from p2igan_b200 import build_generator
    torch.manual_seed(2024)
    G = build_generator(synth.make_cfg(128, 128))
    _check(G.state_dict(), golden["g128"]["fingerprint"])
    assert sum(p.numel() for p in G.parameters() if p.requires_grad) == 25887984


def test_discriminator_state_dict_matches_reference(golden):
    from p2igan_b200 import build_discriminator
    torch.manual_seed(2024)
    D = build_discriminator(synth.make_cfg(32, 32))
    sd = D.state_dict()
    assert len(sd) == 42
    _check(sd, golden["d32"]["fingerprint"])
    assert sum(p.numel() for p in D.parameters()) == 1690884
    assert sd._metadata["d2d.0"]["spectral_norm"] == {"weight.version": 1}


def test_state_dict_roundtrip_and_checkpoint_format(tmp_path):
    from p2igan_b200 import build_discriminator, build_generator
    cfg = synth.make_cfg(32, 32)
    G, D = build_generator(cfg), build_discriminator(cfg)
    path = tmp_path / "latest.pt"
    torch.save({"epoch": 1, "global_step": 3, "generator": G.state_dict(), "discriminator": D.state_dict()}, path)
    ck = torch.load(path, map_location="cpu", weights_only=True)          # scripts/infer.py:183
    G2, D2 = build_generator(cfg), build_discriminator(cfg)
    G2.load_state_dict(ck["generator"])
    D2.load_state_dict(ck["discriminator"])
    for k, v in G.state_dict().items():
        assert torch.equal(v, G2.state_dict()[k])
    for k, v in D.state_dict().items():
        assert torch.equal(v, D2.state_dict()[k])


def test_bad_config_errors():
    import pytest
    from p2igan_b200 import build_generator
    cfg = synth.make_cfg(32, 32, sample_length=20)
    with pytest.raises(ValueError):
        build_generator(cfg)
    with pytest.raises(KeyError):
        build_generator({"model": {"name": "p2igan"}, "data": {}})


def test_modules_deepcopy_and_pickle_roundtrip():
    """EMA-style copies and torch.save(model) must work: transient kernel state (streams, device tables) is not part of the
    pickled module, parameters and buffers are."""
    import copy
    import io
    import synth
    from p2igan_b200 import build_discriminator, build_generator
    cfg = synth.make_cfg(32, 32)
    torch.manual_seed(1)
    for m in (build_generator(cfg), build_discriminator(cfg)):
        c = copy.deepcopy(m)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        r = torch.load(buf, weights_only=False)
        for other in (c, r):
            a, b = m.state_dict(), other.state_dict()
            assert list(a) == list(b)
            assert all(torch.equal(a[k], b[k]) for k in a)
