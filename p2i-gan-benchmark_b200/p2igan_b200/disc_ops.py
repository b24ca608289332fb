"""Discriminator forward / backward as chains of libp2i_sm100a kernels (reference: models/p2igan.py:157-173).

Layer plan for an input [B, T=16, 1, H, W] (the stride-2 convs of the reference run as stride-1 k=2 convs on a
space-to-depth ("s2d") input, which the previous layer's epilogue writes directly):

  2-D branch                                         3-D branch
  pack x -> [B,H,W,64] (16 real channels)            d3d.0  1->32  (1,2,2)  CUDA cores   -> s2d [B,T,H/4,W/4,128]
  d2d.0  16->64  s1  k3  -> s2d [B,H/2,W/2,256]      d3d.2  32->64 (1,2,2)  kt3 k2       -> s2d [B,T,H/8,W/8,256]
  d2d.2  64->128 s2  k2  -> s2d [B,H/4,W/4,512]      d3d.4  64->128 (1,2,2) kt3 k2       -> [B,T,H/8,W/8,128]
  d2d.4  128->256 s2 k2  -> [B,H/4,W/4,256]          d3d.6  128->128 (2,1,1) kt3 k3 st2  -> [B,T/2,H/8,W/8,128]
  d2d.6  256->256 s1 k3  -> [B,H/4,W/4,256]          d3d.8  128->1 1x1x1 + mean_t + bilinear resize   (tail kernel)
  d2d.8  256->1   s1     -> [B,H/4,W/4] f32          fused = sigmoid(alpha2d) * out2d + resized        (tail kernel)
"""
from __future__ import annotations

import ctypes
import struct
from typing import Dict

import torch

from . import _overlap
from ._lib import LIB, ptr, require_cuda, stream

D2D = [0, 2, 4, 6, 8]
D3D = [0, 2, 4, 6, 8]

# (name, Cout, Cin, KT, ksize, s2, cin_pad, keep_t) for the tensor-core layers
TC_LAYERS = [
    ("d2d.0", 64, 16, 1, 3, 0, 64, 0),
    ("d2d.2", 128, 64, 1, 3, 1, 0, 0),
    ("d2d.4", 256, 128, 1, 3, 1, 0, 0),
    ("d2d.6", 256, 256, 1, 3, 0, 256, 0),
    ("d3d.2", 64, 32, 3, 3, 1, 0, 0),
    ("d3d.4", 128, 64, 3, 3, 1, 0, 0),
    ("d3d.6", 128, 128, 3, 3, 0, 128, 1),
]
ALL_SN = ["d2d.0", "d2d.2", "d2d.4", "d2d.6", "d2d.8", "d3d.0", "d3d.2", "d3d.4", "d3d.6", "d3d.8"]


def conv_desc(samples, T_in, T_out, H, W, Cin, Cout, kt, ksize, pad, pad_t, stride_t=1, t_transposed=0, act=0, mask_mode=0,
              out_mode=0):
    """P2iConvDesc as a C int array (16 ints)."""
    vals = [samples, T_in, T_out, H, W, Cin, Cout, kt, ksize, pad, pad_t, stride_t, t_transposed, act, mask_mode, out_mode]
    return (ctypes.c_int * 16)(*vals)


def conv_igemm(x, w, desc, residual=None, mask=None, bias=None, out=None):
    LIB.call("p2i_conv_igemm", ptr(x), ptr(w), desc, ptr(residual), ptr(mask), ptr(bias), ptr(out), stream())
    return out


def conv_wgrad(x, dy, dW, desc):
    LIB.call("p2i_conv_wgrad", ptr(x), ptr(dy), ptr(dW), desc, stream())
    return dW


def _mod(D, name):
    seq, idx = name.split(".")
    return getattr(D, seq)[int(idx)]


N_SETS = 3   # D is called three times per training step (fake, real, fake-for-G): each call owns one operand set


class _OperandSet:
    """Packed bf16 operands, sigma scalars and the (u, v) snapshot of ONE forward call."""

    def __init__(self, mods, dev):
        self.sigma = torch.zeros(len(ALL_SN), dtype=torch.float32, device=dev)
        self.u = {n: torch.empty_like(mods[n].weight_u) for n in ALL_SN}
        self.v = {n: torch.empty_like(mods[n].weight_v) for n in ALL_SN}
        sn = b""
        self.sn_scratch = torch.zeros(len(ALL_SN), 2 + 512, dtype=torch.float32, device=dev)
        self.max_rows = self.max_cols = 1
        for i, n in enumerate(ALL_SN):
            m = mods[n]
            rows = m.weight_orig.shape[0]
            cols = m.weight_orig.numel() // rows
            self.max_rows, self.max_cols = max(self.max_rows, rows), max(self.max_cols, cols)
            sn += struct.pack("<QQQQQQQii", m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(),
                              self.sigma[i:i + 1].data_ptr(), self.u[n].data_ptr(), self.v[n].data_ptr(),
                              self.sn_scratch[i].data_ptr(), rows, cols)
        self.sn_table = torch.frombuffer(bytearray(sn), dtype=torch.uint8).to(dev)
        self.w: Dict[str, torch.Tensor] = {}
        self.wt: Dict[str, torch.Tensor] = {}
        pk = b""
        for name, Cout, Cin, KT, k, s2, cin_pad, keep_t in TC_LAYERS:
            taps = KT * (4 if s2 else k * k)
            cinp = 4 * Cin if s2 else cin_pad
            self.w[name] = torch.zeros(taps, Cout, cinp, dtype=torch.bfloat16, device=dev)
            self.wt[name] = torch.zeros(taps, cinp, Cout, dtype=torch.bfloat16, device=dev)
            si = ALL_SN.index(name)
            pk += struct.pack("<QQQQiiiiiiii", mods[name].weight_orig.data_ptr(), self.sigma[si:si + 1].data_ptr(),
                              self.w[name].data_ptr(), self.wt[name].data_ptr(), Cout, Cin, KT, k, s2, cinp, keep_t, 0)
        self.pack_table = torch.frombuffer(bytearray(pk), dtype=torch.uint8).to(dev)
        self.bwd = None          # lazily built by disc_bwd (persistent gradient buffers + device table)

    def sig(self, name):
        i = ALL_SN.index(name)
        return self.sigma[i:i + 1]


class DiscState:
    """Per-module cache: N_SETS rotating operand sets (so a saved forward context stays valid until its backward ran,
    without cloning anything) and their device tables."""

    def __init__(self, D):
        self.dev = D.alpha2d.device
        self.mods = {n: _mod(D, n) for n in ALL_SN}
        self.key = tuple(m.weight_orig.data_ptr() for m in self.mods.values())
        self.sets = [_OperandSet(self.mods, self.dev) for _ in range(N_SETS)]
        self.calls = 0
        # the 3-D branch runs on a side stream next to the 2-D branch (fork / join with events; CUDA-graph capturable): its
        # thin CUDA-core first layer then overlaps the 2-D branch's tensor-core kernels instead of queueing behind them
        self.side = torch.cuda.Stream(device=self.dev)
        self.aux = torch.cuda.Stream(device=self.dev)      # bias column-sums of the backward pass
        # second lane for the paired fake / real calls of the D update (forward_pair_ctx, disc_bwd.backward_pair)
        self.lane = torch.cuda.Stream(device=self.dev)
        self.side2 = torch.cuda.Stream(device=self.dev)
        self.aux2 = torch.cuda.Stream(device=self.dev)

    def next_set(self) -> _OperandSet:
        s = self.sets[self.calls % N_SETS]
        self.calls += 1
        return s


def _state(D) -> DiscState:
    st = getattr(D, "_p2i_state", None)
    key = tuple(_mod(D, n).weight_orig.data_ptr() for n in ALL_SN)
    if st is None or st.key != key:
        st = DiscState(D)
        D._p2i_state = st
    return st


def _check_input(D, x):
    require_cuda(x)
    B, T, C, H, W = x.shape
    if T * C != D.in_channels or T != 16:
        raise ValueError(f"P2IDiscriminator expects {D.in_channels} frames, got {T * C}")
    if H % 16 or W % 16:
        raise ValueError("P2IDiscriminator requires H and W to be multiples of 16")
    return B, T, H, W


def _prepare_operands(D, state) -> _OperandSet:
    """The spectral-norm hook of ONE forward call on the current stream: in train() mode one power iteration that updates
    weight_u / weight_v in place, sigma, and the packed bf16 GEMM operands W / sigma of the call's operand set."""
    st = state.next_set()
    LIB.call("p2i_spectral_norm", ptr(st.sn_table), len(ALL_SN), st.max_rows, st.max_cols, 1 if D.training else 0, stream())
    LIB.call("p2i_disc_pack_weights", ptr(st.pack_table), len(TC_LAYERS), stream())
    return st


def _alloc_forward(B, T, H, W, dev):
    """Every buffer of one forward call (allocate on the MAIN stream before any fork: the caching allocator is per stream)."""
    bf = torch.bfloat16
    T2 = (T + 2 - 3) // 2 + 1
    return dict(
        a0=torch.empty(B, H, W, 64, dtype=bf, device=dev), y1=torch.empty(B, H // 2, W // 2, 256, dtype=bf, device=dev),
        y2=torch.empty(B, H // 4, W // 4, 512, dtype=bf, device=dev), y3=torch.empty(B, H // 4, W // 4, 256, dtype=bf, device=dev),
        y4=torch.empty(B, H // 4, W // 4, 256, dtype=bf, device=dev), o2d=torch.empty(B, H // 4, W // 4, dtype=torch.float32, device=dev),
        z1=torch.empty(B, T, H // 4, W // 4, 128, dtype=bf, device=dev), z2=torch.empty(B, T, H // 8, W // 8, 256, dtype=bf, device=dev),
        z3=torch.empty(B, T, H // 8, W // 8, 128, dtype=bf, device=dev), z4=torch.empty(B, T2, H // 8, W // 8, 128, dtype=bf, device=dev),
        m=torch.empty(B, H // 8, W // 8, dtype=torch.float32, device=dev),
        fused=torch.empty(B, (H // 4) * (W // 4), dtype=torch.float32, device=dev), T2=T2)


def _forward_body(D, state, st, xf, bufs, side, save):
    """The convolution stacks and the fused tail of one forward call on the CURRENT stream (2-D branch) and `side` (3-D branch)."""
    B, T, H, W = xf.shape
    mods = state.mods
    bias = {n: mods[n].bias.detach() for n in ALL_SN}
    a0, y1, y2, y3, y4, o2d = (bufs[k] for k in ("a0", "y1", "y2", "y3", "y4", "o2d"))
    z1, z2, z3, z4, m, fused, T2 = (bufs[k] for k in ("z1", "z2", "z3", "z4", "m", "fused", "T2"))
    main = torch.cuda.current_stream()
    side = _overlap.pick(side, main, _overlap.D_BRANCH)
    side.wait_stream(main)
    # ---- 3-D branch (side stream)
    with torch.cuda.stream(side):
        LIB.call("p2i_d3d_first_fwd", ptr(xf), ptr(mods["d3d.0"].weight_orig.detach()), ptr(st.sig("d3d.0")), ptr(bias["d3d.0"]),
                 ptr(z1), B, T, H, W, stream())
        conv_igemm(z1, st.w["d3d.2"], conv_desc(B, T, T, H // 4, W // 4, 128, 64, 3, 2, 1, 1, act=2, out_mode=1), bias=bias["d3d.2"], out=z2)
        conv_igemm(z2, st.w["d3d.4"], conv_desc(B, T, T, H // 8, W // 8, 256, 128, 3, 2, 1, 1, act=2), bias=bias["d3d.4"], out=z3)
        conv_igemm(z3, st.w["d3d.6"], conv_desc(B, T, T2, H // 8, W // 8, 128, 128, 3, 3, 1, 1, stride_t=2, act=2), bias=bias["d3d.6"], out=z4)
    # ---- 2-D branch (current stream)
    LIB.call("p2i_disc_pack_input", ptr(xf), ptr(a0), B, 16, H, W, stream())
    conv_igemm(a0, st.w["d2d.0"], conv_desc(B, 1, 1, H, W, 64, 64, 1, 3, 1, 0, act=2, out_mode=1), bias=bias["d2d.0"], out=y1)
    conv_igemm(y1, st.w["d2d.2"], conv_desc(B, 1, 1, H // 2, W // 2, 256, 128, 1, 2, 1, 0, act=2, out_mode=1), bias=bias["d2d.2"], out=y2)
    conv_igemm(y2, st.w["d2d.4"], conv_desc(B, 1, 1, H // 4, W // 4, 512, 256, 1, 2, 1, 0, act=2), bias=bias["d2d.4"], out=y3)
    conv_igemm(y3, st.w["d2d.6"], conv_desc(B, 1, 1, H // 4, W // 4, 256, 256, 1, 3, 1, 0, act=2), bias=bias["d2d.6"], out=y4)
    LIB.call("p2i_d2d_last_fwd", ptr(y4), ptr(mods["d2d.8"].weight_orig.detach()), ptr(st.sig("d2d.8")), ptr(bias["d2d.8"]),
             ptr(o2d), B, H // 4, W // 4, 256, stream())
    main.wait_stream(side)
    # ---- tail
    LIB.call("p2i_disc_tail_fwd", ptr(z4), ptr(mods["d3d.8"].weight_orig.detach()), ptr(st.sig("d3d.8")), ptr(bias["d3d.8"]),
             ptr(o2d), ptr(D.alpha2d.detach()), ptr(m), ptr(fused), B, T2, H // 8, W // 8, 128, H // 4, W // 4, stream())
    ctx = None
    if save:
        ctx = dict(xf=xf, a0=a0, y1=y1, y2=y2, y3=y3, y4=y4, o2d=o2d, z1=z1, z2=z2, z3=z3, z4=z4, m=m, dims=(B, T, H, W, T2),
                   set=st)
    return fused, ctx


def forward_ctx(D, x: torch.Tensor, save: bool = False):
    """x [B,T,1,H,W] f32 -> (logits [B,(H/4)(W/4)] f32, ctx).  Runs the spectral-norm hook semantics first:
    in train() mode every call performs one power iteration and updates weight_u / weight_v in place."""
    B, T, H, W = _check_input(D, x)
    state = _state(D)
    xf = x.detach().reshape(B, T, H, W).contiguous().float()
    st = _prepare_operands(D, state)
    bufs = _alloc_forward(B, T, H, W, x.device)
    return _forward_body(D, state, st, xf, bufs, state.side, save)


def forward_pair_ctx(D, xa: torch.Tensor, xb: torch.Tensor, save: bool = False):
    """D(xa) then D(xb) -- the two discriminator calls of the D update (scripts/train.py:264-265) -- with the reference's
    semantics (two successive power iterations: call A uses sigma after the first, call B after the second) but with the two
    convolution stacks on two stream lanes, so that one call's CUDA-core kernels (thin first 3-D layer, weight packing, tail)
    run under the other call's tensor-core kernels.  -> ((logits_a, ctx_a), (logits_b, ctx_b))."""
    Ba, T, H, W = _check_input(D, xa)
    Bb, _, _, _ = _check_input(D, xb)
    state = _state(D)
    dev = xa.device
    xfa = xa.detach().reshape(Ba, T, H, W).contiguous().float()
    xfb = xb.detach().reshape(Bb, T, H, W).contiguous().float()
    bufs_a, bufs_b = _alloc_forward(Ba, T, H, W, dev), _alloc_forward(Bb, T, H, W, dev)
    main = torch.cuda.current_stream()
    lane = _overlap.pick(state.lane, main, _overlap.D_PAIR)
    st_a = _prepare_operands(D, state)
    lane.wait_stream(main)
    with torch.cuda.stream(lane):
        out_a = _forward_body(D, state, st_a, xfa, bufs_a, state.side, save)
    st_b = _prepare_operands(D, state)             # second power iteration: after the first (same stream), next to call A's convs
    out_b = _forward_body(D, state, st_b, xfb, bufs_b, state.side2, save)
    main.wait_stream(lane)
    return out_a, out_b


def discriminator_forward_pair(D, xa, xb):
    """(D(xa), D(xb)) as one autograd node with the two calls on concurrent stream lanes (same values as two calls)."""
    from .disc_bwd import DiscriminatorPairFn
    return DiscriminatorPairFn.apply(D, xa, xb, *[p for _, p in D.named_parameters()])


def discriminator_forward(D, x):
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in D.parameters()))
    if needs_grad:
        from .disc_bwd import DiscriminatorFn
        names = [n for n, _ in D.named_parameters()]
        return DiscriminatorFn.apply(D, x, *[p for _, p in D.named_parameters()])
    return forward_ctx(D, x)[0]
