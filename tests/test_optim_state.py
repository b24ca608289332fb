"""FusedAdam <-> torch.optim.Adam state_dict compatibility for the reference's checkpoint format (scripts/train.py:125-136,
475-485): the discriminator optimiser holds all 22 parameters incl. the never-used alpha3d, so `optimizer_d`
written by the reference loads here and ours loads there.  Host logic only (no kernel launch)."""
import torch

import synth
from p2igan_b200 import FusedAdam, build_discriminator


def _ref_adam_after_one_step(D):
    opt = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.99))
    g = torch.Generator().manual_seed(0)
    for n, p in D.named_parameters():
        if n != "alpha3d":                       # alpha3d never receives a gradient (models/p2igan.py:145)
            p.grad = torch.randn(p.shape, generator=g)
    opt.step()
    return opt


def test_reference_adam_state_loads_into_fused_adam_and_back():
    torch.manual_seed(1)
    D = build_discriminator(synth.make_cfg(32, 32))
    ref = _ref_adam_after_one_step(D)
    sd = ref.state_dict()
    assert len(sd["param_groups"][0]["params"]) == 22 and 1 not in sd["state"]          # alpha3d: index 1, no state

    ours = FusedAdam(list(D.parameters()), lr=1e-4, betas=(0.0, 0.99))
    ours.load_state_dict(sd)                    # raised ValueError (21 vs 22 params) before the fix
    params = list(D.parameters())
    for i, st in sd["state"].items():
        mine = ours.state[params[i]]
        assert torch.equal(mine["exp_avg"], st["exp_avg"]) and torch.equal(mine["exp_avg_sq"], st["exp_avg_sq"])
        assert float(mine["step"]) == 1.0
    back = ours.state_dict()
    assert sorted(back["state"].keys()) == sorted(sd["state"].keys())
    steps = [st["step"] for st in back["state"].values()]
    assert all(s.device.type == "cpu" and s.dim() == 0 for s in steps)
    assert len({s.data_ptr() for s in steps}) == len(steps)                              # independent scalars
    fresh = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.0, 0.99))
    fresh.load_state_dict(back)
    for p in D.parameters():
        if p.grad is not None:
            p.grad.mul_(0.5)
    fresh.step()                                # the reference's optimiser continues from our checkpoint
    assert all(float(st["step"]) == 2.0 for st in fresh.state.values())


def test_train_step_builds_the_reference_param_groups():
    """GANTrainStep's optimisers mirror Adam(G.parameters()) / Adam(D.parameters()): same number of entries per group."""
    import inspect
    from p2igan_b200 import train_step
    src = inspect.getsource(train_step.GANTrainStep.__init__)
    assert "FusedAdam(list(discriminator.parameters())" in src
