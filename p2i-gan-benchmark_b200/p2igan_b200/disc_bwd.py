"""Backward of the discriminator (chain of libp2i_sm100a kernels; see disc_ops.py for the layer plan)."""
from __future__ import annotations

import struct

import torch

from . import _overlap
from ._lib import LIB, ptr, stream
from .disc_ops import ALL_SN, TC_LAYERS, _state, conv_desc, conv_igemm, conv_wgrad, forward_ctx


class _BwdBuffers:
    """Persistent gradient buffers of one operand set + the device table of the spectral-norm backward."""

    def __init__(self, mods, st, dev):
        f32 = torch.float32
        tc = {n: (Cout, Cin, KT, k, s2, (4 * Cin if s2 else cinp)) for n, Cout, Cin, KT, k, s2, cinp, _ in TC_LAYERS}
        sizes = {}
        for n in ALL_SN:
            sizes[n] = st.w[n].numel() if n in tc else mods[n].weight_orig.numel()
        nb = {n: mods[n].bias.numel() for n in ALL_SN}
        total = sum(sizes.values()) + len(ALL_SN)
        self.arena = torch.zeros(total, dtype=f32, device=dev)       # scratch, zeroed once per backward (single memset)
        o = 0
        self.G = {}
        for n in ALL_SN:
            self.G[n] = self.arena[o:o + sizes[n]]; o += sizes[n]
        self.inner = self.arena[o:o + len(ALL_SN)]; o += len(ALL_SN)
        self.mods, self.st, self.tc = mods, st, tc
        self.key, self.table = None, None

    def table_for(self, targets, dev):
        """Device table of the spectral-norm backward writing (+=) into ``targets[name + '.weight_orig']``."""
        key = tuple(targets[n + ".weight_orig"].data_ptr() for n in ALL_SN)
        if key == self.key:
            return self.table
        mods, st, tc = self.mods, self.st, self.tc
        tab = b""
        for n in ALL_SN:
            m = mods[n]
            if n in tc:
                Cout, Cin, KT, k, s2, cinp = tc[n]
                packed = 1
            else:
                Cout, Cin = m.weight_orig.shape[0], m.weight_orig.shape[1]
                per = m.weight_orig.numel() // (Cout * Cin)
                KT, k, s2, cinp, packed = (3, 3, 0, Cin, 0) if per == 27 else ((1, 3, 0, Cin, 0) if per == 9 else (1, 1, 0, Cin, 0))
            tab += struct.pack("<QQQQQQiiiiiiii", self.G[n].data_ptr(), m.weight_orig.data_ptr(), st.u[n].data_ptr(),
                               st.v[n].data_ptr(), st.sig(n).data_ptr(), targets[n + ".weight_orig"].data_ptr(), Cout, Cin, KT,
                               k, s2, cinp, packed, 0)
        self.table = torch.frombuffer(bytearray(tab), dtype=torch.uint8).to(dev)
        self.key = key
        return self.table


def prepare(D):
    """Allocate the persistent backward buffers of every operand set (call once before CUDA-graph capture)."""
    state = _state(D)
    for st in state.sets:
        if st.bwd is None:
            st.bwd = _BwdBuffers(state.mods, st, state.dev)


def _colsum(g, out):
    C = g.shape[-1]
    LIB.call("p2i_colsum_bf16", ptr(g), ptr(out), g.numel() // C, C, stream())


def backward(D, ctx, dfused, need_params: bool, need_input: bool):
    """dfused f32 [B, (H/4)(W/4)] -> ({param name: grad}, dx f32 [B,T,1,H,W] or None)."""
    state = _state(D)
    st = ctx["set"]
    mods = state.mods
    B, T, H, W, T2 = ctx["dims"]
    dev = dfused.device
    bf, f32 = torch.bfloat16, torch.float32
    w, wt = st.w, st.wt
    if st.bwd is None:
        st.bwd = _BwdBuffers(mods, st, dev)
    bb = st.bwd
    bb.arena.zero_()
    G = bb.G if need_params else {}
    # gradient targets: preallocated .grad (flat-gradient mode, accumulated in place) or fresh zero buffers for autograd
    tg, fresh = {}, {}
    if need_params:
        for n, p in D.named_parameters():
            if n == "alpha3d" or not p.requires_grad:
                continue
            if p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == f32:
                tg[n] = p.grad
            else:
                fresh[n] = torch.zeros_like(p, dtype=f32, memory_format=torch.contiguous_format)
                tg[n] = fresh[n]
    dbias = {n: tg[n + ".bias"] for n in ALL_SN} if need_params else {}
    dfused = dfused.detach().contiguous().float()
    h4, w4, h8, w8 = H // 4, W // 4, H // 8, W // 8

    # ---- tail
    d_o2d = torch.empty(B, h4, w4, dtype=f32, device=dev)
    dpre_z4 = torch.empty_like(ctx["z4"])
    LIB.call("p2i_disc_tail_bwd", ptr(dfused), ptr(ctx["o2d"]), ptr(D.alpha2d.detach()), ptr(ctx["z4"]),
             ptr(mods["d3d.8"].weight_orig.detach()), ptr(st.sig("d3d.8")), ptr(d_o2d), ptr(tg.get("alpha2d")),
             ptr(dpre_z4), ptr(G.get("d3d.8")), ptr(dbias.get("d3d.8")), B, T2, h8, w8, 128, h4, w4, stream())

    def tc_layer(name, x_in, dpre, fdesc, ddesc, mask, out):
        """wgrad + bias grad (if needed) and dgrad of one tensor-core layer -> d(pre-activation) of the layer below."""
        if need_params:
            # the bias gradient (HBM-bound column sum of dpre) goes to the aux stream, next to this layer's wgrad / dgrad GEMMs
            cur = torch.cuda.current_stream()
            aux.wait_stream(cur)
            with torch.cuda.stream(aux):
                _colsum(dpre, dbias[name])
            waux.wait_stream(cur)
            with torch.cuda.stream(waux):        # weight gradient: off the critical path (only the spectral-norm backward reads it)
                conv_wgrad(x_in, dpre, G[name], fdesc)
        if ddesc is None:
            return None
        conv_igemm(dpre, wt[name], ddesc, mask=mask, out=out)
        return out

    # buffers of both branches are allocated on the main stream before the fork (per-stream caching allocator)
    dpre_y4 = torch.empty_like(ctx["y4"])
    e2 = [torch.empty_like(ctx["y3"]), torch.empty(B, H // 2, W // 2, 128, dtype=bf, device=dev), torch.empty(B, H, W, 64, dtype=bf, device=dev),
          torch.empty(B, H, W, 64, dtype=bf, device=dev) if need_input else None]
    e3 = [torch.empty_like(ctx["z3"]), torch.empty(B, T, h4, w4, 64, dtype=bf, device=dev), torch.empty(B, T, H // 2, W // 2, 32, dtype=bf, device=dev)]
    dx = torch.empty(B, T, H, W, dtype=f32, device=dev) if need_input else None
    main = torch.cuda.current_stream()
    side, aux = _overlap.pick(state.side, main, _overlap.D_BRANCH), _overlap.pick(state.aux, main, _overlap.D_COLSUM)
    waux = _overlap.pick(state.aux, main, _overlap.D_WGRAD)
    side.wait_stream(main)

    # ---- 3-D branch (top-down) on the side stream: its CUDA-core first-layer kernels overlap the 2-D branch's GEMMs
    with torch.cuda.stream(side):
        d = tc_layer("d3d.6", ctx["z3"], dpre_z4, conv_desc(B, T, T2, h8, w8, 128, 128, 3, 3, 1, 1, stride_t=2),
                     conv_desc(B, T2, T, h8, w8, 128, 128, 3, 3, 1, 1, stride_t=2, t_transposed=1, mask_mode=2), ctx["z3"], e3[0])
        d = tc_layer("d3d.4", ctx["z2"], d, conv_desc(B, T, T, h8, w8, 256, 128, 3, 2, 1, 1),
                     conv_desc(B, T, T, h8, w8, 128, 256, 3, 2, 0, 1, mask_mode=2, out_mode=2), ctx["z2"], e3[1])
        d = tc_layer("d3d.2", ctx["z1"], d, conv_desc(B, T, T, h4, w4, 128, 64, 3, 2, 1, 1),
                     conv_desc(B, T, T, h4, w4, 64, 128, 3, 2, 0, 1, mask_mode=2, out_mode=2), ctx["z1"], e3[2])
        LIB.call("p2i_d3d_first_bwd", ptr(d), ptr(ctx["xf"]), ptr(mods["d3d.0"].weight_orig.detach()), ptr(st.sig("d3d.0")),
                 ptr(G.get("d3d.0")), ptr(dbias.get("d3d.0")), ptr(dx), B, T, H, W, stream())

    # ---- d2d.8 and the 2-D branch (top-down) on the main stream
    LIB.call("p2i_d2d_last_bwd", ptr(d_o2d), ptr(ctx["y4"]), ptr(mods["d2d.8"].weight_orig.detach()), ptr(st.sig("d2d.8")),
             ptr(dpre_y4), ptr(G.get("d2d.8")), ptr(dbias.get("d2d.8")), B, h4, w4, 256, stream())
    d = tc_layer("d2d.6", ctx["y3"], dpre_y4, conv_desc(B, 1, 1, h4, w4, 256, 256, 1, 3, 1, 0),
                 conv_desc(B, 1, 1, h4, w4, 256, 256, 1, 3, 1, 0, mask_mode=2), ctx["y3"], e2[0])
    d = tc_layer("d2d.4", ctx["y2"], d, conv_desc(B, 1, 1, h4, w4, 512, 256, 1, 2, 1, 0),
                 conv_desc(B, 1, 1, h4, w4, 256, 512, 1, 2, 0, 0, mask_mode=2, out_mode=2), ctx["y2"], e2[1])
    d = tc_layer("d2d.2", ctx["y1"], d, conv_desc(B, 1, 1, H // 2, W // 2, 256, 128, 1, 2, 1, 0),
                 conv_desc(B, 1, 1, H // 2, W // 2, 128, 256, 1, 2, 0, 0, mask_mode=2, out_mode=2), ctx["y1"], e2[2])
    d_a0 = tc_layer("d2d.0", ctx["a0"], d, conv_desc(B, 1, 1, H, W, 64, 64, 1, 3, 1, 0),
                    conv_desc(B, 1, 1, H, W, 64, 64, 1, 3, 1, 0) if need_input else None, None, e2[3])
    main.wait_stream(side)
    if need_params:                 # aux was forked in this call (joining a stream that is not part of a capture is an error)
        main.wait_stream(aux)
        main.wait_stream(waux)
    if need_input:
        LIB.call("p2i_disc_unpack_input_grad", ptr(d_a0), ptr(dx), B, 16, H, W, stream())
        dx = dx.view(B, T, 1, H, W)

    # ---- spectral-norm backward (+ un-pack of tensor-core gradients): two launches for all 10 layers
    if need_params:
        LIB.call("p2i_spectral_norm_bwd", ptr(bb.table_for(tg, dev)), len(ALL_SN), ptr(bb.inner), stream())
    return fresh, dx


class DiscriminatorFn(torch.autograd.Function):
    """Whole-discriminator autograd node.  Parameters with a preallocated .grad are accumulated in place by the kernels
    (flat-gradient mode); the others receive their gradient through autograd as usual."""

    @staticmethod
    def forward(ctx, D, x, *params):
        out, saved = forward_ctx(D, x, save=True)
        ctx.D, ctx.saved = D, saved
        ctx.names = [n for n, _ in D.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, dout):
        need_input = ctx.needs_input_grad[1]
        need_params = any(f for f, n in zip(ctx.needs_input_grad[2:], ctx.names) if n != "alpha3d")
        grads, dx = backward(ctx.D, ctx.saved, dout, need_params, need_input)
        ctx.saved = None
        pg = tuple(grads.get(n) if (need_params and ctx.needs_input_grad[2 + i]) else None for i, n in enumerate(ctx.names))
        return (None, dx) + pg
