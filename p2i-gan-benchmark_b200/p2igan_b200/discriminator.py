"""P2IDiscriminator on sm_100a kernels (reference: p2igan_bench/models/p2igan.py:115-173).

Dual-branch patch discriminator: a 2-D branch (T folded into channels) and a 3-D branch, every conv
spectrally normalised, fused as sigmoid(alpha2d) * out2d + resize(mean_t(out3d)).  Drop-in: same
constructor, sub-module names and the 42 ``state_dict`` keys (``weight_orig/_u/_v``, ``bias``,
``alpha2d``, ``alpha3d``).  torch's ``spectral_norm`` is used ONLY as the parameter container (so that
keys, metadata and the RNG stream of initialisation are identical); its forward hook never runs --
the power iteration, the normalisation and the convolutions are our kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import BaseNetwork, C2, C3

D2D_LAYERS = [(None, 64, 1), (64, 128, 2), (128, 256, 2), (256, 256, 1), (256, 1, 1)]        # (cin, cout, stride)
D3D_LAYERS = [(1, 32, (1, 2, 2)), (32, 64, (1, 2, 2)), (64, 128, (1, 2, 2)), (128, 128, (2, 1, 1))]


class P2IDiscriminator(BaseNetwork):
    def __init__(self, in_channels: int = 16, init_weights: bool = True):
        super().__init__()
        self.in_channels = in_channels
        seq = []
        for i, (cin, cout, s) in enumerate(D2D_LAYERS):
            seq.append(C2(in_channels if cin is None else cin, cout, k=3, s=s, p=1))
            if i < len(D2D_LAYERS) - 1:
                seq.append(nn.LeakyReLU(0.2, True))
        self.d2d = nn.Sequential(*seq)
        seq = []
        for cin, cout, st in D3D_LAYERS:
            seq.append(C3(cin, cout, kt=3, ks=3, st=st, pt=(1, 1, 1)))
            seq.append(nn.LeakyReLU(0.2, True))
        seq.append(nn.utils.spectral_norm(nn.Conv3d(128, 1, kernel_size=1)))
        self.d3d = nn.Sequential(*seq)
        self.alpha2d = nn.Parameter(torch.tensor(0.0))
        self.alpha3d = nn.Parameter(torch.tensor(0.0))
        if init_weights:
            self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                nn.init.kaiming_normal_(m.weight, a=0.2, nonlinearity="leaky_relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        from . import disc_ops
        return disc_ops.discriminator_forward(self, x)

    def forward_pair(self, xa, xb):
        """(self(xa), self(xb)) -- identical values and state updates (two successive power iterations in train mode) -- with
        the two calls on concurrent stream lanes: the fake / real pair of the D update (scripts/train.py:264-265)."""
        from . import disc_ops
        return disc_ops.discriminator_forward_pair(self, xa, xb)

    def __getstate__(self):
        """copy.deepcopy / pickling of the module (EMA copies, torch.save(model)): drop the per-device kernel state (operand
        sets, device tables, CUDA streams); it is rebuilt on the next forward."""
        d = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        d = dict(d)
        d.pop("_p2i_state", None)
        return d
