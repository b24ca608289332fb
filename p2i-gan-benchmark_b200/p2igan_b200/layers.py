"""Host-side mirror of p2igan_bench/modules/layer.py + deconv_pytorch.py for the P2I-GAN hot path.

Same class names, constructor arguments, parameter names/shapes and registration order as the
reference, so ``state_dict`` keys and the RNG stream of random initialisation match.
The arithmetic is NOT PyTorch's: every ``forward`` calls the sm_100a kernels in libp2i_sm100a.so.
Modules accept/return the reference's NCHW float32 tensors; inside the trunk the generator keeps
activations in NHWC bf16 and calls the ``*_cl`` methods directly.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
from torch.nn import init

from . import ops
from ._lib import pack_do_table, require_cuda


class BaseNetwork(nn.Module):
    """Weight-initialisation helper (reference: layer.py:14-40)."""

    def init_weights(self, init_type: str = "kaiming", gain: float = 0.02):
        def visit(m):
            name = type(m).__name__
            if "BatchNorm2d" in name:
                init.normal_(m.weight.data, 1.0, gain)
                init.constant_(m.bias.data, 0.0)
                return
            if not hasattr(m, "weight") or not ("Conv" in name or "Linear" in name):
                return
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=gain)
            else:
                raise NotImplementedError(f"initialization method [{init_type}] is not implemented")
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)

        self.apply(visit)


class DOConv2d(nn.Module):
    """Over-parameterised conv (reference: deconv_pytorch.py:13-132).  Parameters: W [Cout, Cin/g, k*k],
    and for k*k > 1: D [Cin, k*k, k*k] (zeros), D_diag [Cin, k*k, k*k] (identity, frozen but a Parameter)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, D_mul=None, stride=1, padding=1, dilation=1, groups=1,
                 bias=False, padding_mode="zeros", simam=False):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError("in_channels must be divisible by groups")
        if out_channels % groups != 0:
            raise ValueError("out_channels must be divisible by groups")
        if padding_mode != "zeros" or dilation != 1 or stride != 1 or simam or bias:
            raise ValueError("p2igan_b200.DOConv2d supports the configuration P2I-GAN uses: stride 1, zero padding, "
                             "no dilation/bias/simam")
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation = (stride, stride), (padding, padding), (dilation, dilation)
        mn = kernel_size * kernel_size
        self.D_mul = mn if D_mul is None or mn <= 1 else D_mul
        if self.D_mul != mn:
            raise ValueError("D_mul != k*k is not used by P2I-GAN and is not supported")
        self.W = nn.Parameter(torch.empty(out_channels, in_channels // groups, self.D_mul))
        init.kaiming_uniform_(self.W, a=math.sqrt(5))
        if mn > 1:
            self.D = nn.Parameter(torch.zeros(in_channels, mn, self.D_mul))
            eye = torch.eye(mn, dtype=torch.float32).reshape(1, mn, mn)
            self.D_diag = nn.Parameter(eye.repeat(in_channels, 1, 1), requires_grad=False)
        self.register_parameter("bias", None)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, groups={self.groups}"

    def composed_weight_cl(self) -> torch.Tensor:
        """bf16 [k*k, Cout, Cin] operand for a single groups=1 layer (per-layer path; the generator batches this)."""
        C = self.in_channels
        if self.groups != 1 or self.out_channels != C or self.kernel_size != (3, 3):
            raise ValueError("composed_weight_cl: only square groups=1 3x3 layers")
        out = torch.empty(9, C, C, dtype=torch.bfloat16, device=self.W.device)
        tab = pack_do_table([(self.W.data_ptr(), self.D.data_ptr(), self.D_diag.data_ptr(), out.data_ptr(), 0, C)])
        tab_dev = torch.frombuffer(bytearray(tab), dtype=torch.uint8).to(self.W.device)
        ops.doconv_compose(tab_dev, 1, C)
        return out

    def forward_cl(self, x_cl, residual=None, relu=False):
        return ops.conv2d_cl(x_cl, self.composed_weight_cl(), residual, relu)

    def forward(self, x):
        require_cuda(x)
        return ops.from_cl(self.forward_cl(ops.to_cl(x.contiguous().float())))


class BasicConv_do(nn.Module):
    """DOConv2d (+ ReLU)  (reference: layer.py:68-94)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, bias=False, norm=False, relu=True, transpose=False,
                 relu_method=nn.ReLU, groups=1, norm_method=nn.BatchNorm2d):
        super().__init__()
        if transpose or norm or (relu and relu_method is not nn.ReLU):
            raise ValueError("p2igan_b200.BasicConv_do: transpose/norm/non-ReLU variants are unused by P2I-GAN")
        layers: List[nn.Module] = [DOConv2d(in_channel, out_channel, kernel_size, padding=kernel_size // 2, stride=stride,
                                            bias=bias, groups=groups)]
        if relu:
            layers.append(nn.ReLU(inplace=True))
        self.relu = relu
        self.main = nn.Sequential(*layers)

    def forward_cl(self, x_cl, residual=None):
        return self.main[0].forward_cl(x_cl, residual, self.relu)

    def forward(self, x):
        require_cuda(x)
        return ops.from_cl(self.forward_cl(ops.to_cl(x.contiguous().float())))


class ResBlock_do(nn.Module):
    """conv-ReLU-conv + identity  (reference: layer.py:126-135)."""

    def __init__(self, out_channel):
        super().__init__()
        self.main = nn.Sequential(
            BasicConv_do(out_channel, out_channel, kernel_size=3, stride=1, relu=True),
            BasicConv_do(out_channel, out_channel, kernel_size=3, stride=1, relu=False),
        )

    def forward_cl(self, x_cl):
        return self.main[1].forward_cl(self.main[0].forward_cl(x_cl), residual=x_cl)

    def forward(self, x):
        require_cuda(x)
        return ops.from_cl(self.forward_cl(ops.to_cl(x.contiguous().float())))


class EBlock(nn.Module):
    """num_res residual blocks  (reference: models/p2igan.py:176-183)."""

    def __init__(self, out_channel, num_res=8, ResBlock=ResBlock_do):
        super().__init__()
        self.layers = nn.Sequential(*[ResBlock(out_channel) for _ in range(num_res)])

    def forward_cl(self, x_cl):
        for blk in self.layers:
            x_cl = blk.forward_cl(x_cl)
        return x_cl

    def forward(self, x):
        require_cuda(x)
        return ops.from_cl(self.forward_cl(ops.to_cl(x.contiguous().float())))


class DownsampleDuplicateChannels(nn.Module):
    """max-pool 2x2 + duplicate every channel (reference: layer.py:200-214).  Parameter-free; the
    generator fuses the three levels it needs into one kernel (ops.pyramid_fwd); the module's own forward runs one level."""

    def __init__(self, length):
        super().__init__()
        self.t = length

    def forward(self, x):
        """x [B,C,H,W] f32 -> [B,2C,H/2,W/2] (one level; differentiable).  Same checks as the reference (:207-208)."""
        require_cuda(x)
        b, c, h, w = x.shape
        if c % self.t != 0:
            raise ValueError(f"Channel dimension {c} must be divisible by length {self.t}.")
        return _DownsampleDupFn.apply(x.contiguous().float())


class _DownsampleDupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.downsample_dup_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.downsample_dup_bwd(x, dy.contiguous().float())


class AttentionBlock(nn.Module):
    """Parameter container for the per-pixel gate relu(x + x*conv1x1(x))  (reference: layer.py:296-304)."""

    def __init__(self, c=1):
        super().__init__()
        self.conv = nn.Conv1d(c, c, kernel_size=1)


class InputBlock(nn.Module):
    """Gate x depth + 4-NN inverse-distance interpolation of the observed points (reference: layer.py:307-361)."""

    def __init__(self, depth=4, k=4, rho=2.0, tau=0.05, chunk=1000):
        super().__init__()
        self.layers = nn.ModuleList([AttentionBlock(16) for _ in range(depth)])
        if depth != 2 or k != 4 or abs(rho - 2.0) > 1e-6:
            raise ValueError("p2igan_b200.InputBlock implements the P2I-GAN configuration depth=2, k=4, rho=2")
        self.k, self.rho, self.tau, self.chunk = k, rho, tau, chunk
        self._cache = None          # ops.IdwTableCache: sample 0's neighbour table across calls (static gauge masks)

    def gate_params(self):
        c0, c1 = self.layers[0].conv, self.layers[1].conv
        return c0.weight, c0.bias, c1.weight, c1.bias

    def forward_ctx(self, inp: torch.Tensor, mask: torch.Tensor, save_for_backward: bool = False):
        """Returns (out [B,16,H,W] f32, ctx) where ctx holds what the backward needs."""
        require_cuda(inp, mask)
        B, D, H, W = inp.shape
        inp = inp.contiguous().float()
        mask = mask.contiguous().float()
        cap = D * H * W
        pts, counts, src = ops.points_extract(mask, cap)
        w0, b0, w1, b1 = (p.detach().contiguous() for p in self.gate_params())
        vals, _ = ops.gate_points_fwd(inp, pts, counts, w0, b0, w1, b1)
        key = (D, H, W, float(self.tau), cap, str(inp.device))
        if self._cache is None or self._cache.key != key:
            self._cache = ops.IdwTableCache((D, H, W), self.tau, cap, inp.device)
        self._cache.check(pts, counts)
        out, table = ops.idw_knn_fwd(pts, vals, counts, src, (D, H, W), self.tau, cache=self._cache)
        ctx = (inp, pts, counts, src, table) if save_for_backward else None
        return out, ctx

    def forward(self, inp, mask):
        return self.forward_ctx(inp, mask)[0]


class UPPos(nn.Module):
    """bilinear x2 (align_corners) -> *(1 + 2*sigmoid(pos) - 1) -> 1x1 conv + bias -> ReLU (reference: layer.py:384-399).
    Computed as: 1x1 projection at LOW resolution on the tensor cores, then one fused upsample/modulate/bias/ReLU
    kernel (the projection commutes with the per-channel upsample and the per-pixel scale)."""

    def __init__(self, in_ch, out_ch, T, H, W):
        super().__init__()
        self.T = T
        self.pos = nn.Parameter(torch.zeros(1, 1, H, W))
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.proj = nn.Conv2d(in_ch, out_ch, kernel_size=1, bias=True)

    def proj_weight_cl(self):
        w = self.proj.weight.detach()
        return w.reshape(1, w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous()

    def forward_cl(self, x_cl, skip=None, w_cl=None):
        z = ops.conv2d_cl(x_cl, w_cl if w_cl is not None else self.proj_weight_cl())
        return ops.upmod_fwd(z, self.pos.detach().reshape(self.pos.shape[-2], self.pos.shape[-1]).contiguous(),
                             self.proj.bias.detach().contiguous(), skip)

    def forward(self, x):
        require_cuda(x)
        return ops.from_cl(self.forward_cl(ops.to_cl(x.contiguous().float())))


def C2(cin, cout, k=3, s=1, p=1):
    return nn.utils.spectral_norm(nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p))


def C3(cin, cout, kt=3, ks=3, st=(1, 1, 1), pt=(1, 1, 1)):
    return nn.utils.spectral_norm(nn.Conv3d(cin, cout, kernel_size=(kt, ks, ks), stride=st, padding=pt))
