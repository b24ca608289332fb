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

    def _init_state(self, gi, group):
        plist = [p for p in group["params"] if p.requires_grad or p.grad is not None]
        dev = plist[0].device
        # one device buffer per group: [step, lr/(1-b1^t), 1/sqrt(1-b2^t), pad]; every state["step"] is a 0-dim view of [0]
        buf = self._stepbufs.get(gi)
        if buf is None or buf.device != dev:
            buf = torch.zeros(4, dtype=torch.float32, device=dev)
            self._stepbufs[gi] = buf
            seeded = False
        else:
            seeded = True
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

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plist)
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
