"""Backward of the discriminator (chain of libp2i_sm100a kernels; see disc_ops.py for the layer plan)."""
from __future__ import annotations

import struct

import torch

from ._lib import LIB, ptr, stream
from .disc_ops import ALL_SN, TC_LAYERS, _state, conv_desc, conv_igemm, conv_wgrad, forward_ctx


def _colsum(g, C):
    out = torch.zeros(C, dtype=torch.float32, device=g.device)
    LIB.call("p2i_colsum_bf16", ptr(g), ptr(out), g.numel() // C, C, stream())
    return out


def backward(D, ctx, dfused, need_params: bool, need_input: bool):
    """dfused f32 [B, (H/4)(W/4)] -> ({param name: grad}, dx f32 [B,T,1,H,W] or None)."""
    st = _state(D)
    B, T, H, W, T2 = ctx["dims"]
    dev = dfused.device
    bf, f32 = torch.bfloat16, torch.float32
    mods = st.mods
    sig = ctx["sigma"]
    w, wt = ctx["w"], ctx["wt"]

    def sg(name):
        i = ALL_SN.index(name)
        return sig[i:i + 1]

    grads = {}
    G = {}                                   # dL/dW_sn per layer (packed or plain)
    dfused = dfused.detach().contiguous().float()
    h4, w4, h8, w8 = H // 4, W // 4, H // 8, W // 8

    # ---- tail
    d_o2d = torch.empty(B, h4, w4, dtype=f32, device=dev)
    dalpha = torch.zeros((), dtype=f32, device=dev)
    dpre_z4 = torch.empty_like(ctx["z4"])
    if need_params:
        G["d3d.8"] = torch.zeros(128, dtype=f32, device=dev)
        grads["d3d.8.bias"] = torch.zeros(1, dtype=f32, device=dev)
    LIB.call("p2i_disc_tail_bwd", ptr(dfused), ptr(ctx["o2d"]), ptr(D.alpha2d.detach()), ptr(ctx["z4"]),
             ptr(mods["d3d.8"].weight_orig.detach()), ptr(sg("d3d.8")), ptr(d_o2d), ptr(dalpha), ptr(dpre_z4),
             ptr(G.get("d3d.8")), ptr(grads.get("d3d.8.bias")), B, T2, h8, w8, 128, h4, w4, stream())
    grads["alpha2d"] = dalpha

    # ---- d2d.8
    dpre_y4 = torch.empty_like(ctx["y4"])
    if need_params:
        G["d2d.8"] = torch.zeros(256 * 9, dtype=f32, device=dev)
        grads["d2d.8.bias"] = torch.zeros(1, dtype=f32, device=dev)
    LIB.call("p2i_d2d_last_bwd", ptr(d_o2d), ptr(ctx["y4"]), ptr(mods["d2d.8"].weight_orig.detach()), ptr(sg("d2d.8")),
             ptr(dpre_y4), ptr(G.get("d2d.8")), ptr(grads.get("d2d.8.bias")), B, h4, w4, 256, stream())

    def tc_layer(name, x_in, dpre, fdesc, ddesc, mask, out_shape):
        """wgrad + bias grad (if needed) and dgrad of one tensor-core layer. Returns d(pre-activation) of the layer below."""
        if need_params:
            G[name] = torch.zeros(w[name].shape, dtype=f32, device=dev)
            conv_wgrad(x_in, dpre, G[name], fdesc)
            grads[name + ".bias"] = _colsum(dpre, dpre.shape[-1])
        if ddesc is None:
            return None
        out = torch.empty(out_shape, dtype=bf, device=dev)
        conv_igemm(dpre, wt[name], ddesc, mask=mask, out=out)
        return out

    # ---- 2-D branch (top-down)
    d = tc_layer("d2d.6", ctx["y3"], dpre_y4, conv_desc(B, 1, 1, h4, w4, 256, 256, 1, 3, 1, 0),
                 conv_desc(B, 1, 1, h4, w4, 256, 256, 1, 3, 1, 0, mask_mode=2), ctx["y3"], ctx["y3"].shape)
    d = tc_layer("d2d.4", ctx["y2"], d, conv_desc(B, 1, 1, h4, w4, 512, 256, 1, 2, 1, 0),
                 conv_desc(B, 1, 1, h4, w4, 256, 512, 1, 2, 0, 0, mask_mode=2, out_mode=2), ctx["y2"], (B, H // 2, W // 2, 128))
    d = tc_layer("d2d.2", ctx["y1"], d, conv_desc(B, 1, 1, H // 2, W // 2, 256, 128, 1, 2, 1, 0),
                 conv_desc(B, 1, 1, H // 2, W // 2, 128, 256, 1, 2, 0, 0, mask_mode=2, out_mode=2), ctx["y1"], (B, H, W, 64))
    d_a0 = tc_layer("d2d.0", ctx["a0"], d, conv_desc(B, 1, 1, H, W, 64, 64, 1, 3, 1, 0),
                    conv_desc(B, 1, 1, H, W, 64, 64, 1, 3, 1, 0) if need_input else None, None, (B, H, W, 64))

    # ---- 3-D branch (top-down)
    d = tc_layer("d3d.6", ctx["z3"], dpre_z4, conv_desc(B, T, T2, h8, w8, 128, 128, 3, 3, 1, 1, stride_t=2),
                 conv_desc(B, T2, T, h8, w8, 128, 128, 3, 3, 1, 1, stride_t=2, t_transposed=1, mask_mode=2), ctx["z3"], ctx["z3"].shape)
    d = tc_layer("d3d.4", ctx["z2"], d, conv_desc(B, T, T, h8, w8, 256, 128, 3, 2, 1, 1),
                 conv_desc(B, T, T, h8, w8, 128, 256, 3, 2, 0, 1, mask_mode=2, out_mode=2), ctx["z2"], (B, T, h4, w4, 64))
    d = tc_layer("d3d.2", ctx["z1"], d, conv_desc(B, T, T, h4, w4, 128, 64, 3, 2, 1, 1),
                 conv_desc(B, T, T, h4, w4, 64, 128, 3, 2, 0, 1, mask_mode=2, out_mode=2), ctx["z1"], (B, T, H // 2, W // 2, 32))
    dx = torch.empty(B, T, H, W, dtype=f32, device=dev) if need_input else None
    if need_params:
        G["d3d.0"] = torch.zeros(32 * 27, dtype=f32, device=dev)
        grads["d3d.0.bias"] = torch.zeros(32, dtype=f32, device=dev)
    LIB.call("p2i_d3d_first_bwd", ptr(d), ptr(ctx["xf"]), ptr(mods["d3d.0"].weight_orig.detach()), ptr(sg("d3d.0")),
             ptr(G.get("d3d.0")), ptr(grads.get("d3d.0.bias")), ptr(dx), B, T, H, W, stream())
    if need_input:
        LIB.call("p2i_disc_unpack_input_grad", ptr(d_a0), ptr(dx), B, 16, H, W, stream())
        dx = dx.view(B, T, 1, H, W)

    # ---- spectral-norm backward (+ un-pack of tensor-core gradients), one launch for all 10 layers
    if need_params:
        tc = {n: (Cout, Cin, KT, k, s2, cinp if not s2 else 4 * Cin) for n, Cout, Cin, KT, k, s2, cinp, _ in TC_LAYERS}
        tab = b""
        outs = {}
        for n in ALL_SN:
            m = mods[n]
            dW = torch.empty_like(m.weight_orig)
            outs[n] = dW
            if n in tc:
                Cout, Cin, KT, k, s2, cinp = tc[n]
                packed = 1
            else:
                Cout, Cin = m.weight_orig.shape[0], m.weight_orig.shape[1]
                per = m.weight_orig.numel() // (Cout * Cin)
                KT, k, s2, cinp, packed = (3, 3, 0, Cin, 0) if per == 27 else ((1, 3, 0, Cin, 0) if per == 9 else (1, 1, 0, Cin, 0))
            tab += struct.pack("<QQQQQQiiiiiiii", G[n].data_ptr(), m.weight_orig.data_ptr(), ctx["u"][n].data_ptr(),
                               ctx["v"][n].data_ptr(), sg(n).data_ptr(), dW.data_ptr(), Cout, Cin, KT, k, s2, cinp, packed, 0)
        tab_dev = torch.frombuffer(bytearray(tab), dtype=torch.uint8).to(dev)
        LIB.call("p2i_spectral_norm_bwd", ptr(tab_dev), len(ALL_SN), stream())
        for n in ALL_SN:
            grads[n + ".weight_orig"] = outs[n]
    return grads, dx


class DiscriminatorFn(torch.autograd.Function):
    """Whole-discriminator autograd node."""

    @staticmethod
    def forward(ctx, D, x, *params):
        out, saved = forward_ctx(D, x, save=True)
        ctx.D, ctx.saved = D, saved
        ctx.names = [n for n, _ in D.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, dout):
        need_input = ctx.needs_input_grad[1]
        need_params = any(ctx.needs_input_grad[2:])
        grads, dx = backward(ctx.D, ctx.saved, dout, need_params, need_input)
        ctx.saved = None
        pg = tuple(grads.get(n) if (need_params and ctx.needs_input_grad[2 + i]) else None for i, n in enumerate(ctx.names))
        return (None, dx) + pg
