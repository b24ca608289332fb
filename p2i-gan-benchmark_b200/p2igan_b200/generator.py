"""P2IGenerator on sm_100a kernels (reference: p2igan_bench/models/p2igan.py:23-112).

Drop-in: same constructor signature, sub-module names and ``state_dict`` layout (113 keys).  ``forward``
runs the whole trunk in NHWC bf16 through libp2i_sm100a.so:

    InputBlock (ordered point extraction, gates at the observed points, exact 4-NN IDW)      CUDA cores
    one batched DO-Conv weight composition for the 32 ResBlock convs (bf16 GEMM operand)      CUDA cores
    Convsin stem + repeat_interleave, max-pool/duplicate pyramid (x4, x8)                     CUDA cores, HBM bound
    4 levels x 4 ResBlocks: 3x3 convs as tcgen05 implicit GEMM, ReLU / +x fused               tensor cores
    3 x UPPos: 1x1 projection (tcgen05) at low resolution + fused upsample/modulate/ReLU(+skip)
    ConvsOut grouped 1x1 + tanh -> [B,T,1,H,W] float32
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from . import _overlap
from ._lib import pack_do_grad_table, pack_do_table, require_cuda
from .layers import (BaseNetwork, BasicConv_do, DownsampleDuplicateChannels, EBlock, InputBlock, ResBlock_do, UPPos)


class P2IGenerator(BaseNetwork):
    def __init__(self, config, length: int = 16, num_res: int = 4, inference: bool = False, init_weights: bool = True):
        super().__init__()
        data_cfg = config.get("data_loader") or config["data"]["train"]
        self.keep = data_cfg.get("mask", {}).get("keep", 0)
        self.H, self.W = data_cfg["h"], data_cfg["w"]
        length = data_cfg.get("sample_length", length)
        if length != 16:
            # the reference hard-wires AttentionBlock(16) and base_channel = 4*T
            raise ValueError(f"P2IGenerator requires sample_length == 16 (got {length}); longer events are handled by "
                             "16-frame sliding windows (scripts/infer.py:217-241)")
        if self.H % 8 or self.W % 8:
            raise ValueError("P2IGenerator requires H and W to be multiples of 8")
        self.length = length
        self.inference = inference            # the reference's `_eval` variants are never selected by its scripts
        self.num_res = num_res

        self.input = InputBlock(depth=2, k=4, rho=2.0, tau=0.05, chunk=16384)
        base = 64
        self.Decoder = nn.ModuleList([EBlock(base * m, num_res, ResBlock=ResBlock_do) for m in (1, 2, 4, 8)])
        self.ConvsOut = nn.ModuleList([BasicConv_do(base, length, kernel_size=1, relu=False, stride=1, groups=4)])
        self.UP = nn.ModuleList([
            UPPos(in_ch=base * 2, out_ch=base, H=self.H, W=self.W, T=length),
            UPPos(in_ch=base * 4, out_ch=base * 2, H=self.H // 2, W=self.W // 2, T=length),
            UPPos(in_ch=base * 8, out_ch=base * 4, H=self.H // 4, W=self.W // 4, T=length),
        ])
        self.Convsin = nn.ModuleList([BasicConv_do(length, base, kernel_size=3, relu=False, stride=1, groups=4)])
        self.downsample = DownsampleDuplicateChannels(length=length)
        if init_weights:
            self.init_weights()
        self._wcache: Optional[Dict] = None

    # ------------------------------------------------------------------ composed-weight cache
    def _res_convs(self):
        for level in range(4):
            for blk in self.Decoder[level].layers:
                yield level, blk.main[0].main[0]
                yield level, blk.main[1].main[0]

    def _weights(self, need_dgrad: bool = False) -> Dict:
        """bf16 GEMM operands for all convs; recomposed only when a parameter changed (version counters)."""
        convs = list(self._res_convs())
        dev = convs[0][1].W.device
        ptr_key = tuple(c.W.data_ptr() for _, c in convs) + tuple(c.D.data_ptr() for _, c in convs)
        ver_key = tuple(c.W._version for _, c in convs) + tuple(c.D._version for _, c in convs) + \
            tuple(u.proj.weight._version for u in self.UP) + (self.Convsin[0].main[0].W._version,
                                                             self.Convsin[0].main[0].D._version)
        wc = self._wcache
        if wc is None or wc["ptr_key"] != ptr_key or (need_dgrad and wc["bufs_t"][0] is None):
            bufs = [torch.empty(9, c.in_channels, c.in_channels, dtype=torch.bfloat16, device=dev) for _, c in convs]
            bufs_t = [torch.empty_like(b) for b in bufs] if need_dgrad else [None] * len(bufs)
            tab = pack_do_table([(c.W.data_ptr(), c.D.data_ptr(), c.D_diag.data_ptr(), b.data_ptr(),
                                  bt.data_ptr() if bt is not None else 0, c.in_channels)
                                 for (_, c), b, bt in zip(convs, bufs, bufs_t)])
            wc = {"ptr_key": ptr_key, "ver_key": None, "bufs": bufs, "bufs_t": bufs_t,
                  "table": torch.frombuffer(bytearray(tab), dtype=torch.uint8).to(dev)}
            self._wcache = wc
        if wc["ver_key"] != ver_key or self.training:      # training: weights change every step (also under CUDA graphs)
            ops.doconv_compose(wc["table"], len(convs), max(c.in_channels for _, c in convs))
            s = self.Convsin[0].main[0]
            wc["stem"] = ops.doconv_compose_stem(s.W.detach(), s.D.detach(), s.D_diag.detach())
            wc["up"] = [u.proj_weight_cl() for u in self.UP]
            wc["up_t"] = [w.transpose(1, 2).contiguous() for w in wc["up"]]
            wc["ver_key"] = ver_key
        return wc

    # ------------------------------------------------------------------ forward
    def trunk_cl(self, x_in: torch.Tensor, wc: Dict, keep: Optional[List] = None):
        """Everything after the InputBlock. x_in [B,16,H,W] f32 -> (out [B,16,H,W] f32 post-tanh)."""
        bufs = wc["bufs"]

        def eblock(level, x):
            base = level * 2 * self.num_res
            for r in range(self.num_res):
                y = ops.conv2d_cl(x, bufs[base + 2 * r], None, True)
                x = ops.conv2d_cl(y, bufs[base + 2 * r + 1], x, False)
            return x

        stem = ops.stem_fwd(x_in, wc["stem"])
        x4, x8 = ops.pyramid_fwd(stem)
        r = eblock(3, x8)
        r = self.UP[2].forward_cl(r, skip=x4, w_cl=wc["up"][2])
        r = eblock(2, r)
        r = self.UP[1].forward_cl(r, w_cl=wc["up"][1])
        r = eblock(1, r)
        r = self.UP[0].forward_cl(r, w_cl=wc["up"][0])
        r = eblock(0, r)
        w_out = self.ConvsOut[0].main[0].W.detach().reshape(16, 16).contiguous()
        return ops.head_fwd(r, w_out)

    def forward(self, masked_frames, masks):
        require_cuda(masked_frames, masks)
        b, t, c, h, w = masked_frames.shape
        if (t * c, h, w) != (16, self.H, self.W):
            raise ValueError(f"P2IGenerator built for 16x{self.H}x{self.W}, got {t * c}x{h}x{w}")
        mf = masked_frames.reshape(b, c * t, h, w)
        mk = masks.reshape(b, c * t, h, w)
        params = [p for _, p in self.named_parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            out = _GeneratorFn.apply(self, mf, mk, *params)
        else:
            x = self.input(mf, mk)
            out = self.trunk_cl(x, self._weights())
        return out.view(b, t, c, h, w)

    # ------------------------------------------------------------------ training path
    def _grad_table(self, wc: Dict):
        """One fp32 wgrad arena for the 32 DO-Conv layers + the device table of the composition backward."""
        convs = list(self._res_convs())
        key = wc["ptr_key"]
        gt = getattr(self, "_gtable", None)
        if gt is None or gt["key"] != key:
            dev = convs[0][1].W.device
            sizes = [9 * c.in_channels * c.in_channels for _, c in convs]
            arena = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            views, o = [], 0
            for n, (_, c) in zip(sizes, convs):
                views.append(arena[o:o + n].view(9, c.in_channels, c.in_channels))
                o += n
            gt = {"key": key, "arena": arena, "views": views}
            self._gtable = gt
        return gt

    def __getstate__(self):
        """copy.deepcopy / pickling of the module: drop transient kernel state (operand caches, gradient arena, CUDA stream)."""
        d = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        d = dict(d)
        for k in ("_wcache", "_gtable", "_side", "_bucket_hook"):
            d.pop(k, None)
        d["_wcache"] = None
        return d

    def _side_stream(self, device, which: int = 0) -> torch.cuda.Stream:
        """0: InputBlock / weight gradients of levels 0-2; 1: stem weight gradient; 2: weight gradients of level 3, HIGH priority
        (its CTAs are placed before those of the lower levels' backlog when SMs free up -- see _backward_train)."""
        sts = getattr(self, "_side", None)
        if sts is None or sts[0].device != torch.device(device):
            sts = [torch.cuda.Stream(device=device), torch.cuda.Stream(device=device), torch.cuda.Stream(device=device, priority=-1)]
            self._side = sts
        return sts[which]

    def _forward_train(self, mf, mk):
        """Forward that keeps what the backward needs. Returns (out f32 [B,16,H,W], saved dict)."""
        # the InputBlock (latency-bound CUDA-core kernels) runs on a side stream next to the weight composition
        main = torch.cuda.current_stream()
        side = _overlap.pick(self._side_stream(mf.device), main, _overlap.G_INPUT)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            x_in, ictx = self.input.forward_ctx(mf, mk, save_for_backward=True)
        wc = self._weights(need_dgrad=True)
        bufs = wc["bufs"]
        main.wait_stream(side)
        sv = {"ictx": ictx, "x_in": x_in, "wc": wc, "res": {}, "up": {}}

        def eblock(level, x):
            base = level * 2 * self.num_res
            acts = []
            for r in range(self.num_res):
                y = ops.conv2d_cl(x, bufs[base + 2 * r], None, True)
                o = ops.conv2d_cl(y, bufs[base + 2 * r + 1], x, False)
                acts.append((x, y))
                x = o
            sv["res"][level] = acts
            return x

        def up(i, x, skip=None):
            U = self.UP[i]
            z = ops.conv2d_cl(x, wc["up"][i])
            pos = U.pos.detach().reshape(U.pos.shape[-2], U.pos.shape[-1]).contiguous()
            bias = U.proj.bias.detach().contiguous()
            sv["up"][i] = (x, z, pos, bias)
            return ops.upmod_fwd(z, pos, bias, skip)

        stem = ops.stem_fwd(x_in, wc["stem"])
        x4, x8 = ops.pyramid_fwd(stem)
        sv["stem"] = stem
        r = eblock(3, x8)
        r = up(2, r, skip=x4)
        r = eblock(2, r)
        r = up(1, r)
        r = eblock(1, r)
        r = up(0, r)
        r = eblock(0, r)
        w_out = self.ConvsOut[0].main[0].W.detach().reshape(16, 16).contiguous()
        out = ops.head_fwd(r, w_out)
        sv["r"], sv["w_out"], sv["out"] = r, w_out, out
        return out, sv

    def _grad_targets(self):
        """{name: tensor to accumulate the gradient into}.  Parameters whose .grad is preallocated (flat-gradient
        mode of the trainer) are accumulated in place; the others get a fresh zero buffer that is handed to autograd."""
        tg, fresh = {}, {}
        for n, p in self.named_parameters():
            if not p.requires_grad:
                continue
            if p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32:
                tg[n] = p.grad
            else:
                fresh[n] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                tg[n] = fresh[n]
        return tg, fresh

    def _backward_train(self, sv, dout):
        """dout f32 [B,16,H,W] -> {parameter name: gradient tensor for autograd} (empty in flat-gradient mode)."""
        wc = sv["wc"]
        bufs_t = wc["bufs_t"]
        gt = self._grad_table(wc)
        gt["arena"].zero_()
        gviews = gt["views"]
        tg, fresh = self._grad_targets()
        d, dw_out = ops.head_bwd(dout.contiguous(), sv["out"], sv["r"], sv["w_out"])
        tg["ConvsOut.0.main.0.W"].add_(dw_out.reshape(16, 16, 1))

        # Weight gradients are off the critical path (they only feed the arena): they run on the side stream while the
        # data-gradient chain and the CUDA-core glue (UPPos / pyramid / stem / InputBlock backward) continue on the main
        # stream.  Every tensor a side-stream kernel reads is kept alive in `keep` until the join (the caching allocator is
        # per stream: a block freed on the main stream could otherwise be reused while the side stream still reads it).
        main = torch.cuda.current_stream()
        wside = _overlap.pick(self._side_stream(dout.device), main, _overlap.G_WGRAD)
        keep = []

        # Level 3 (512 channels) is reached last but owns 74 % of the gradient bytes: its weight gradients go to a HIGH-priority
        # stream so that they overtake the backlog of the lower levels' weight gradients, its buckets complete right behind its
        # data-gradient chain and their exchange / Adam update overlap the backlog instead of trailing it.
        wside3 = _overlap.pick(self._side_stream(dout.device, 2), wside, _overlap.G_WGRAD3)
        used = set()

        def wgrad_side(x, g, ks, out, st=None):
            st = wside if st is None else st
            keep.extend((x, g))
            st.wait_stream(main)
            used.add(st)
            with torch.cuda.stream(st):
                ops.conv2d_wgrad(x, g, ks, out=out)

        # The DO-Conv composition backward (arena -> dW, dD) runs per BUCKET of layers, on the stream of the weight gradients,
        # as soon as the bucket's last weight gradient is launched; `_bucket_hook` (set by the data-parallel step) then
        # exchanges that bucket's range of the flat gradient buffer while the backward pass continues.
        convs = list(self._res_convs())
        buckets = self._bwd_buckets()
        hook = getattr(self, "_bucket_hook", None)

        def finish_bucket(bi):
            level, blocks = buckets[bi]
            idx = [level * 2 * self.num_res + 2 * r + j for r in sorted(blocks) for j in (0, 1)]
            names = [f"Decoder.{level}.layers.{r}.main.{j}.main.0" for r in sorted(blocks) for j in (0, 1)]
            key = tuple(tg[n + ".W"].data_ptr() for n in names) + tuple(tg[n + ".D"].data_ptr() for n in names)
            tabs = gt.setdefault("bwd_tables", {})
            if bi not in tabs or tabs[bi][0] != key:
                tab = pack_do_grad_table([(convs[i][1].W.data_ptr(), convs[i][1].D.data_ptr(), convs[i][1].D_diag.data_ptr(),
                                           gviews[i].data_ptr(), tg[n + ".W"].data_ptr(), tg[n + ".D"].data_ptr(),
                                           convs[i][1].in_channels) for i, n in zip(idx, names)])
                tabs[bi] = (key, torch.frombuffer(bytearray(tab), dtype=torch.uint8).to(dout.device))
            st = wside3 if level == 3 else wside
            with torch.cuda.stream(st):
                ops.doconv_compose_bwd(tabs[bi][1], len(idx), convs[idx[0]][1].in_channels)
            if hook is not None:
                hook([n + sfx for n in names for sfx in (".W", ".D")], st)

        def eblock_bwd(level, d):
            base = level * 2 * self.num_res
            for r in reversed(range(self.num_res)):
                a, y = sv["res"][level][r]
                st = wside3 if level == 3 else None
                wgrad_side(y, d, 3, gviews[base + 2 * r + 1], st)
                dy = ops.conv2d_cl(d, bufs_t[base + 2 * r + 1], None, False, mask=y)
                wgrad_side(a, dy, 3, gviews[base + 2 * r], st)
                d = ops.conv2d_cl(dy, bufs_t[base + 2 * r], d, False)
                for bi, (lv, blocks) in enumerate(buckets):
                    if lv == level and min(blocks) == r:        # blocks are visited in descending order: r closes the bucket
                        finish_bucket(bi)
            return d

        def up_bwd(i, d):
            x, z, pos, bias = sv["up"][i]
            dz = ops.upmod_bwd(z, pos, bias, d, tg[f"UP.{i}.proj.bias"], tg[f"UP.{i}.pos"])
            wv = tg[f"UP.{i}.proj.weight"]
            wgrad_side(x, dz, 1, wv.view(1, wv.shape[0], wv.shape[1]))
            return ops.conv2d_cl(dz, wc["up_t"][i])

        d = eblock_bwd(0, d)
        d = up_bwd(0, d)
        d = eblock_bwd(1, d)
        d = up_bwd(1, d)
        d = eblock_bwd(2, d)
        d_x4 = d
        d = up_bwd(2, d)
        d_x8 = eblock_bwd(3, d)
        d_stem = ops.pyramid_bwd(sv["stem"], d_x4, d_x8)
        # The stem's weight gradient (and its composition backward) only feed the gradient buffer: off the critical path, on
        # the side stream behind the last weight gradients, next to the InputBlock backward chain.
        s = self.Convsin[0].main[0]
        dw_stem = torch.zeros(64, 4, 9, dtype=torch.float32, device=dout.device)        # allocated on the main stream
        sside = _overlap.pick(self._side_stream(dout.device, 1), main, _overlap.G_STEMDW)   # NOT the weight-gradient stream: that one has a backlog
        dx_in = ops.stem_bwd_dx(d_stem, sv["x_in"], wc["stem"])
        sside.wait_stream(main)
        with torch.cuda.stream(sside):
            ops.stem_bwd_dw(d_stem, sv["x_in"], wc["stem"], dw_stem)
            ops.doconv_compose_stem_bwd(s.W.detach(), s.D.detach(), s.D_diag.detach(), dw_stem, tg["Convsin.0.main.0.W"],
                                        tg["Convsin.0.main.0.D"])
        keep.extend((d_stem, dw_stem))
        # InputBlock
        inp, pts, counts, src, table = sv["ictx"]
        dvals = ops.idw_knn_bwd(dx_in, table, counts, src, pts.shape[1])
        w0, b0, w1, b1 = (p.detach().contiguous() for p in self.input.gate_params())
        ops.gate_points_bwd(inp, pts, counts, w0, b0, w1, b1, dvals, tg["input.layers.0.conv.weight"],
                            tg["input.layers.0.conv.bias"], tg["input.layers.1.conv.weight"], tg["input.layers.1.conv.bias"])
        for st in used | {wside, sside}:
            if st is not main:
                main.wait_stream(st)
        return fresh

    # (level, residual blocks) per gradient bucket, in the order the backward pass completes them.  The 512-channel level
    # holds 74 % of the gradient bytes and arrives last: one bucket per residual block, so that its exchange keeps pace with
    # the data-gradient chain instead of starting after it.
    def _bwd_buckets(self):
        n = self.num_res
        if n != 4:
            return [(lv, tuple(range(n))) for lv in range(4)]
        return [(0, (0, 1, 2, 3)), (1, (0, 1, 2, 3)), (2, (2, 3)), (2, (0, 1)), (3, (3,)), (3, (2,)), (3, (1,)), (3, (0,))]


class _GeneratorFn(torch.autograd.Function):
    """Whole-generator autograd node: forward and backward are chains of libp2i_sm100a kernels."""

    @staticmethod
    def forward(ctx, gen, mf, mk, *params):
        out, sv = gen._forward_train(mf, mk)
        ctx.gen, ctx.sv = gen, sv
        ctx.names = [n for n, _ in gen.named_parameters()]
        return out

    @staticmethod
    def backward(ctx, dout):
        grads = ctx.gen._backward_train(ctx.sv, dout)
        ctx.sv = None
        return (None, None, None) + tuple(grads.get(n) for n in ctx.names)
