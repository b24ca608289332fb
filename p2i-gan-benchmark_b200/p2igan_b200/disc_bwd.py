"""Backward of the discriminator (chain of libp2i_sm100a kernels; see disc_ops.py for the layer plan)."""
from __future__ import annotations

import struct

import torch

from . import _overlap
from ._lib import LIB, ptr, stream
from .disc_ops import ALL_SN, TC_LAYERS, _state, conv_desc, conv_igemm, conv_wgrad, forward_ctx, forward_pair_ctx


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


def _targets(D, need_params):
    """Gradient targets: preallocated .grad (flat-gradient mode, accumulated in place) or fresh zero buffers for autograd."""
    tg, fresh = {}, {}
    if need_params:
        for n, p in D.named_parameters():
            if n == "alpha3d" or not p.requires_grad:
                continue
            if p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32:
                tg[n] = p.grad
            else:
                fresh[n] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                tg[n] = fresh[n]
    return tg, fresh


def _prepare(D, ctx, dfused, need_params, need_input, tg):
    """Everything one backward call allocates or zeroes -- on the CURRENT (main) stream, before any fork."""
    state = _state(D)
    st = ctx["set"]
    B, T, H, W, T2 = ctx["dims"]
    dev = dfused.device
    bf, f32 = torch.bfloat16, torch.float32
    if st.bwd is None:
        st.bwd = _BwdBuffers(state.mods, st, dev)
    st.bwd.arena.zero_()
    h4, w4 = H // 4, W // 4
    return dict(
        ctx=ctx, st=st, tg=tg, need_params=need_params, need_input=need_input, dfused=dfused.detach().contiguous().float(),
        d_o2d=torch.empty(B, h4, w4, dtype=f32, device=dev), dpre_z4=torch.empty_like(ctx["z4"]), dpre_y4=torch.empty_like(ctx["y4"]),
        e2=[torch.empty_like(ctx["y3"]), torch.empty(B, H // 2, W // 2, 128, dtype=bf, device=dev),
            torch.empty(B, H, W, 64, dtype=bf, device=dev), torch.empty(B, H, W, 64, dtype=bf, device=dev) if need_input else None],
        e3=[torch.empty_like(ctx["z3"]), torch.empty(B, T, h4, w4, 64, dtype=bf, device=dev),
            torch.empty(B, T, H // 2, W // 2, 32, dtype=bf, device=dev)],
        dx=torch.empty(B, T, H, W, dtype=f32, device=dev) if need_input else None)


def _body(D, job, side_stream, aux_stream):
    """All kernels of one backward call except the spectral-norm backward, on the CURRENT stream (2-D branch), `side_stream`
    (3-D branch) and `aux_stream` (bias column sums, weight gradients).  -> dx f32 [B,T,1,H,W] or None."""
    state = _state(D)
    ctx, st, tg = job["ctx"], job["st"], job["tg"]
    need_params, need_input = job["need_params"], job["need_input"]
    mods = state.mods
    B, T, H, W, T2 = ctx["dims"]
    wt = st.wt
    bb = st.bwd
    G = bb.G if need_params else {}
    dbias = {n: tg[n + ".bias"] for n in ALL_SN} if need_params else {}
    h4, w4, h8, w8 = H // 4, W // 4, H // 8, W // 8
    d_o2d, dpre_z4, dpre_y4, e2, e3, dx = (job[k] for k in ("d_o2d", "dpre_z4", "dpre_y4", "e2", "e3", "dx"))

    # ---- tail
    LIB.call("p2i_disc_tail_bwd", ptr(job["dfused"]), ptr(ctx["o2d"]), ptr(D.alpha2d.detach()), ptr(ctx["z4"]),
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

    main = torch.cuda.current_stream()
    side, aux = _overlap.pick(side_stream, main, _overlap.D_BRANCH), _overlap.pick(aux_stream, main, _overlap.D_COLSUM)
    waux = _overlap.pick(aux_stream, main, _overlap.D_WGRAD)
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

    # ---- d2d.8 and the 2-D branch (top-down) on the current stream
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
    return dx


def _sn_backward(D, job):
    """Spectral-norm backward (+ un-pack of tensor-core gradients) of one call: two launches for all 10 layers.  Adds (+=, not
    atomically) into the gradient targets: calls that share targets must be serialised."""
    if job["need_params"]:
        bb = job["st"].bwd
        LIB.call("p2i_spectral_norm_bwd", ptr(bb.table_for(job["tg"], job["dfused"].device)), len(ALL_SN), ptr(bb.inner), stream())


def backward(D, ctx, dfused, need_params: bool, need_input: bool):
    """dfused f32 [B, (H/4)(W/4)] -> ({param name: grad}, dx f32 [B,T,1,H,W] or None)."""
    state = _state(D)
    tg, fresh = _targets(D, need_params)
    job = _prepare(D, ctx, dfused, need_params, need_input, tg)
    dx = _body(D, job, state.side, state.aux)
    _sn_backward(D, job)
    return fresh, dx


def backward_pair(D, ctx_a, ctx_b, da, db, need_params: bool, need_input_a: bool, need_input_b: bool):
    """Backward of forward_pair_ctx: the two calls' chains on two stream lanes (disjoint operand sets and scratch arenas; the
    shared bias / alpha gradients are accumulated with atomics), the two spectral-norm backwards one after the other."""
    state = _state(D)
    tg, fresh = _targets(D, need_params)
    job_a = _prepare(D, ctx_a, da, need_params, need_input_a, tg)
    job_b = _prepare(D, ctx_b, db, need_params, need_input_b, tg)
    main = torch.cuda.current_stream()
    lane = _overlap.pick(state.lane, main, _overlap.D_PAIR)
    lane.wait_stream(main)
    with torch.cuda.stream(lane):
        dxa = _body(D, job_a, state.side, state.aux)
        _sn_backward(D, job_a)
    dxb = _body(D, job_b, state.side2, state.aux2)
    main.wait_stream(lane)
    _sn_backward(D, job_b)
    return fresh, dxa, dxb


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


class DiscriminatorPairFn(torch.autograd.Function):
    """(D(xa), D(xb)) of the D update (scripts/train.py:264-265) as ONE autograd node: same values as two DiscriminatorFn
    calls, but the two forward chains -- and, in backward, the two backward chains -- run on two concurrent stream lanes."""

    @staticmethod
    def forward(ctx, D, xa, xb, *params):
        (out_a, saved_a), (out_b, saved_b) = forward_pair_ctx(D, xa, xb, save=True)
        ctx.D, ctx.saved = D, (saved_a, saved_b)
        ctx.names = [n for n, _ in D.named_parameters()]
        return out_a, out_b

    @staticmethod
    def backward(ctx, da, db):
        need_a, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_params = any(f for f, n in zip(ctx.needs_input_grad[3:], ctx.names) if n != "alpha3d")
        grads, dxa, dxb = backward_pair(ctx.D, ctx.saved[0], ctx.saved[1], da, db, need_params, need_a, need_b)
        ctx.saved = None
        pg = tuple(grads.get(n) if (need_params and ctx.needs_input_grad[3 + i]) else None for i, n in enumerate(ctx.names))
        return (None, dxa, dxb) + pg
