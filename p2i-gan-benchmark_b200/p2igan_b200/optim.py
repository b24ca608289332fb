"""Fused multi-tensor Adam on sm_100a with torch.optim.Adam's state layout (``step``, ``exp_avg``, ``exp_avg_sq``),
so checkpoints written by the reference trainer (scripts/train.py:475-485, ``optimizer_g`` / ``optimizer_d``) load
unchanged.  One kernel launch per optimiser step; the step counter lives on the device (as torch's
``capturable=True`` Adam does), which makes the whole update CUDA-graph capturable."""
from __future__ import annotations

import struct

import torch

from ._lib import LIB, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}
        self._stepbufs = {}
        self._seeded = set()          # groups whose device step counter already reflects the (possibly loaded) state

    def _init_state(self, gi, group):
        # like torch.optim.Adam, state exists only for parameters that have received a gradient (the discriminator's
        # unused `alpha3d`, models/p2igan.py:145, stays in the param group without state -- same indices as the reference)
        plist = [p for p in group["params"] if p.grad is not None]
        dev = plist[0].device
        # one device buffer per group: [step, lr/(1-b1^t), 1/sqrt(1-b2^t), pad]; every state["step"] is a 0-dim view of [0]
        buf = self._stepbufs.get(gi)
        if buf is None or buf.device != dev:
            buf = torch.zeros(4, dtype=torch.float32, device=dev)
            self._stepbufs[gi] = buf
            self._seeded.discard(gi)
        seeded = gi in self._seeded
        step = buf[0]
        for p in plist:
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = step
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif st["step"].data_ptr() != step.data_ptr():     # state loaded from a (torch.optim.Adam) checkpoint
                if not seeded:
                    buf[0] = st["step"].to(dev, torch.float32)
                    seeded = True
                st["step"] = step
        self._seeded.add(gi)

    def state_dict(self):
        """torch.optim.Adam's layout with an independent host scalar per ``step`` (ours are views of one device counter; a
        checkpoint must load into the reference's Adam, scripts/train.py:125-136, whose foreach update increments each
        step tensor separately).  Host synchronisation: checkpoints are written outside the hot loop."""
        sd = super().state_dict()
        host_step = {}                        # one D2H read per device counter
        state = {}
        for k, st in sd["state"].items():     # the packed per-parameter dicts ARE the live ones: copy before editing
            st = dict(st)
            if "step" in st and torch.is_tensor(st["step"]):
                ptr_ = st["step"].data_ptr()
                if ptr_ not in host_step:
                    host_step[ptr_] = st["step"].detach().to("cpu", torch.float32)
                st["step"] = host_step[ptr_].clone()
            state[k] = st
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        """Also valid on an optimiser that has already stepped: the loaded moments replace the tensors the device work table
        points to and the loaded step count replaces the device counter."""
        super().load_state_dict(state_dict)
        self._tables.clear()
        self._seeded.clear()
        for buf in self._stepbufs.values():
            buf.zero_()

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in plist)
        tb = self._tables.get(gi)
        if tb is None or tb[0] != key:
            chunk = LIB.load().p2i_adam_chunk_elems()
            blob, chunks = b"", []
            for ti, p in enumerate(plist):
                st = self.state[p]
                blob += struct.pack("<QQQQq", p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                    st["exp_avg_sq"].data_ptr(), p.numel())
                for c in range((p.numel() + chunk - 1) // chunk):
                    chunks += [ti, c]
            dev = plist[0].device
            tb = (key, torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev),
                  torch.tensor(chunks, dtype=torch.int32, device=dev), len(chunks) // 2)
            self._tables[gi] = tb
        return tb

    # ---- bucketed update: begin_step() once, then apply(params) per gradient bucket (each parameter exactly once per step)
    @torch.no_grad()
    def begin_step(self) -> None:
        """step += 1 and the bias corrections of this optimiser step, on the current stream (before any apply())."""
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            self._init_state(gi, group)
            LIB.call("p2i_adam_tick", ptr(self.state[plist[0]]["step"]), float(group["lr"]), float(group["betas"][0]),
                     float(group["betas"][1]), stream())

    @torch.no_grad()
    def apply(self, params, grad_scale: float = 1.0) -> None:
        """Adam update of `params` (a subset of param group 0 with gradients) on the current stream, using the step counter
        begin_step() advanced.  The work table of every distinct subset is built once and cached."""
        group = self.param_groups[0]
        plist = [p for p in params if p.grad is not None]
        if not plist:
            return
        key = ("sub",) + tuple(id(p) for p in plist)
        _, tens, chunks, n = self._table(key, plist)
        LIB.call("p2i_adam_apply", ptr(tens), ptr(chunks), n, ptr(self.state[plist[0]]["step"]), float(group["lr"]),
                 float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(grad_scale), stream())
        torch.autograd.graph.increment_version(plist)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            self._init_state(gi, group)
            for p in plist:
                if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam needs contiguous float32 parameters and gradients")
            _, tens, chunks, n = self._table(gi, plist)
            LIB.call("p2i_adam_step", ptr(tens), ptr(chunks), n, ptr(self.state[plist[0]]["step"]), float(group["lr"]),
                     float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(grad_scale), stream())
            # the kernel wrote the parameters behind autograd's back: bump their version counters so that
            # version-keyed caches (composed DO-Conv operands) see the update
            torch.autograd.graph.increment_version(plist)
        return loss
