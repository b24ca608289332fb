"""One GAN training iteration in the reference trainer's order (scripts/train.py:240-326), on sm_100a kernels.

    preds = G(masked, masks);  loss_g = ReconstructionLoss(k1)(preds, frames)
    if use_gan:  D(fake.detach()), D(real) -> 0.5*(hinge_real + hinge_fake) -> D backward / Adam
                 freeze D -> adv = gan_loss(D(preds), real, is_disc=False) * adversarial_weight -> loss_g += adv
    G backward / Adam;  unfreeze D

Gradients live in ONE flat fp32 buffer per model (``p.grad`` are views): the backward kernels accumulate into it
directly, ``zero_grad`` is a single memset, the data-parallel exchange is a single NCCL all-reduce(sum) over NVLink
per model (1/world_size folded into the fused Adam kernel), and the Adam work table never changes.  Nothing in the
step synchronises with the host, so the whole iteration can be captured in a CUDA graph (``GraphedStep``).

Data parallelism is new relative to the reference: one process per GPU, each rank holds its own
events; N ranks x b events  ==  one rank x N*b events up to fp32 summation order.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Iterable, List

import torch
import torch.distributed as dist

from .losses import ReconstructionLoss, gan_loss
from .optim import FusedAdam


class FlatGrads:
    """Flat fp32 gradient buffer; ``p.grad`` of every given parameter becomes a view into it.

    ``peer=True`` (multi-GPU, one node): the buffer is a CUDA-IPC allocation mapped by every rank and ``all_reduce`` is
    ONE peer-memory kernel (``peer.PeerAllReduce``, CUDA-graph capturable) instead of an NCCL collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], peer: bool = False, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.peer = None
        if peer:
            from .peer import PeerAllReduce
            self.peer = PeerAllReduce(n, group)
            self.flat = self.peer.tensor[:n]
        else:
            self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        o = 0
        self.offsets = {}                     # id(param) -> (first element, numel) inside the flat buffer
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
            self.offsets[id(p)] = (o, p.numel())
            o += p.numel()
        self.n = o
        self._done = []                       # ranges already exchanged in this step (bucketed mode)
        self._comm = None

    def zero(self):
        self.flat.zero_()

    # ---- bucketed exchange (peer mode): ranges of the buffer are summed over the ranks on a side stream as soon as the
    # backward pass has finalised them; finish() exchanges whatever is left and joins the side stream.
    BUCKET_BLOCKS = int(os.environ.get("P2I_PEER_BLOCKS", "64"))     # CTAs of a bucket exchange: few enough to co-reside with GEMMs

    def comm_stream(self) -> torch.cuda.Stream:
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.flat.device, priority=-1)     # exchange CTAs are placed as soon as a slot frees
        return self._comm

    def exchange_params(self, params, producer: torch.cuda.Stream) -> bool:
        """All-reduce(sum) the contiguous range spanned by `params` (final on stream `producer`) on the comm stream.
        Returns False (nothing launched) when the range is not float4-aligned or not contiguous: finish() covers it."""
        if self.peer is None or self.peer.world == 1:
            return False
        spans = sorted(self.offsets[id(p)] for p in params)
        lo, hi = spans[0][0], spans[-1][0] + spans[-1][1]
        if sum(n for _, n in spans) != hi - lo or lo % 4 or (hi - lo) % 4:
            return False
        comm = self.comm_stream()
        comm.wait_stream(producer)
        with torch.cuda.stream(comm):
            self.peer.all_reduce_range(lo, hi - lo, self.BUCKET_BLOCKS)
        self._done.append((lo, hi))
        return True

    def finish(self, group=None) -> float:
        """Exchange every range no bucket covered, wait for the comm stream; returns the factor that turns the sums into means."""
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        world = dist.get_world_size(group)
        if world == 1:
            return 1.0
        if self.peer is None or not self._done:
            self._done = []
            return self.all_reduce(group)
        main = torch.cuda.current_stream()
        comm = self.comm_stream()
        comm.wait_stream(main)
        done = sorted(self._done)
        rest, pos = [], 0
        n4 = (self.n + 3) // 4 * 4            # the peer buffer is padded to a multiple of 4 elements
        for lo, hi in done + [(n4, n4)]:
            if lo > pos:
                rest.append((pos, lo))
            pos = max(pos, hi)
        with torch.cuda.stream(comm):
            for lo, hi in rest:
                self.peer.all_reduce_range(lo, hi - lo, 0 if hi - lo > (1 << 20) else self.BUCKET_BLOCKS)
        main.wait_stream(comm)
        self._done = []
        return 1.0 / world

    def all_reduce(self, group=None) -> float:
        """Sum over data-parallel ranks (one collective); returns the factor that turns the sum into a mean."""
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        world = dist.get_world_size(group)
        if world == 1:
            return 1.0
        if self.peer is not None:
            self.peer.all_reduce()
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world


class GANTrainStep:
    def __init__(self, cfg: Dict[str, Any], generator, discriminator=None, process_group=None, peer_exchange: bool = False):
        self.cfg = cfg
        self.G, self.D = generator, discriminator
        loss_cfg = cfg.get("loss", {})
        self.use_gan = bool(loss_cfg.get("use_gan", 0)) and discriminator is not None
        self.gan_type = loss_cfg.get("gan_loss", "hinge")
        self.adv_w = loss_cfg.get("adversarial_weight", 0.01)
        self.real_label = loss_cfg.get("target_real_label", 1.0)
        self.fake_label = loss_cfg.get("target_fake_label", 0.0)
        self.rec = ReconstructionLoss(k1_alpha=loss_cfg.get("k1_weight", 0.0))
        oc = cfg["train"]["optimizer"]
        betas = (oc.get("beta1", 0.0), oc.get("beta2", 0.99))
        self.pg = process_group
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.peer_exchange = bool(peer_exchange and multi)
        d_params = [p for n, p in discriminator.named_parameters() if n != "alpha3d"] if self.use_gan else None
        # alpha3d never receives a gradient in the reference (unused Parameter, models/p2igan.py:145): keep grad None
        if self.peer_exchange:
            # CUDA IPC can be unavailable (container policy, no P2P between the devices).  PeerAllReduce takes that decision
            # collectively (it raises on every rank or on none), so all ranks fall back to the NCCL exchange together
            try:
                self.flat_g = FlatGrads(generator.parameters(), peer=True, group=process_group)
                self.flat_d = FlatGrads(d_params, peer=True, group=process_group) if self.use_gan else None
            except RuntimeError as e:
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({e}): using NCCL all-reduce")
                self.peer_exchange = False
        if not self.peer_exchange:
            self.flat_g = FlatGrads(generator.parameters())
            self.flat_d = FlatGrads(d_params) if self.use_gan else None
        self.opt_g = FusedAdam(self.flat_g.params, lr=oc["lr"], betas=betas)
        # Bucketed generator update: the generator reports every finalised bucket of DO-Conv gradients (generator.py); the
        # bucket is exchanged over NVLink right away (peer mode) and its Adam update follows on the same side stream, under
        # the rest of the backward pass.  Not with the NCCL exchange (the gradients are only summed after the backward pass).
        world = dist.get_world_size(process_group) if multi else 1
        self._g_scale = 1.0 / world
        # 1 GPU: always (the bucket's Adam update hides under the backward pass).  N GPUs, peer exchange: opt-in with
        # P2I_BUCKETED=1 -- measured SLOWER at N = 2 (6.52 vs 6.34 ms/step, profiles/r2_bucketed_exchange.txt): level 3 holds 74 % of
        # the bytes and completes last, an exchange that co-resides with the tensor-core CTAs is limited to ~64 registers x 256
        # threads per CTA and no shared memory (100-130 GB/s per 19 MB bucket against 500 GB/s for the full-width kernel), and
        # its CTAs slow the latency-bound kernels they share SMs with.
        self._bucketed = (not multi and os.environ.get("P2I_BUCKETED", "1") != "0") or \
            (multi and self.peer_exchange and os.environ.get("P2I_BUCKETED", "0") == "1")
        self._adam_done = set()
        self._armed = False
        if self._bucketed:
            self._g_named = dict(generator.named_parameters())
            generator._bucket_hook = self._on_bucket
        self.opt_d = None
        if self.use_gan:
            # the optimiser sees ALL discriminator parameters in registration order, alpha3d included (grad None: skipped by
            # the kernel, no state) -- Adam(discriminator.parameters()) of scripts/train.py:131-136, so that `optimizer_d`
            # of a reference checkpoint loads here and ours loads there (22 params in the group, same state indices)
            self.opt_d = FusedAdam(list(discriminator.parameters()), lr=oc["lr"], betas=betas)
            from . import disc_bwd
            disc_bwd.prepare(discriminator)

    def _on_bucket(self, names, producer: torch.cuda.Stream) -> None:
        if not self._armed:          # a backward pass outside step() (validation, user code) must not touch the parameters
            return
        params = [self._g_named[n] for n in names]
        run_on = producer
        if self.peer_exchange:
            if not self.flat_g.exchange_params(params, producer):
                return                                   # left to FlatGrads.finish() and the closing Adam launch
            run_on = self.flat_g.comm_stream()
        with torch.cuda.stream(run_on):
            self.opt_g.apply(params, grad_scale=self._g_scale)
        self._adam_done.update(id(p) for p in params)

    def _gan(self, logits, real, is_disc):
        return gan_loss(logits, real, loss_type=self.gan_type, is_disc=is_disc, target_real_label=self.real_label,
                        target_fake_label=self.fake_label)

    # The iteration is cut at its two data-parallel exchange points so that each piece can be captured in its own
    # CUDA graph with the NCCL all-reduces issued between the replays (``GraphedDPStep``):
    #   seg_a: G fwd, rec loss, D(fake), D(real), D loss, D backward        -> all-reduce(D grads)
    #   seg_b: D Adam, D(G(x)) with D frozen, G loss, G backward             -> all-reduce(G grads)
    #   seg_c: G Adam
    def seg_a(self, frames, masked_frames, masks) -> None:
        G, D = self.G, self.D
        preds = G(masked_frames, masks)
        loss_g, pool, reg = self.rec.tensors(preds, frames)
        self._preds, self._loss_g = preds, loss_g
        self._out = {"rec": loss_g.detach(), "pool": pool, "reg": reg}
        if self.use_gan:
            for p in D.parameters():
                p.requires_grad_(True)
            if hasattr(D, "forward_pair") and os.environ.get("P2I_D_PAIR", "1") != "0":
                logits_fake, logits_real = D.forward_pair(preds.detach(), frames)      # two calls, two stream lanes
            else:
                logits_fake = D(preds.detach())
                logits_real = D(frames)
            loss_d = (self._gan(logits_real, True, True) + self._gan(logits_fake, False, True)) * 0.5
            self.flat_d.zero()
            loss_d.backward()
            self._out["dis"] = loss_d.detach()

    def seg_b(self, d_scale: float = 1.0) -> None:
        D = self.D
        loss_g = self._loss_g
        if self.use_gan:
            self.opt_d.step(grad_scale=d_scale)
            for p in D.parameters():
                p.requires_grad_(False)
            adv = self._gan(D(self._preds), True, False) * self.adv_w
            loss_g = loss_g + adv
            self._out["adv"] = adv.detach()
        self.flat_g.zero()
        if self._bucketed:
            self._adam_done.clear()
            self.opt_g.begin_step()                      # before the backward pass forks its side streams
            self._armed = True
        try:
            loss_g.backward()
        finally:
            self._armed = False
        self._out["total"] = loss_g.detach()
        self._preds = self._loss_g = None

    def seg_c(self, g_scale: float = 1.0) -> None:
        if self._bucketed:
            rest = [p for p in self.flat_g.params if id(p) not in self._adam_done]
            self.opt_g.apply(rest, grad_scale=g_scale)   # whatever no bucket covered (and everything, if no bucket fired)
            self._adam_done.clear()
        else:
            self.opt_g.step(grad_scale=g_scale)
        if self.use_gan:
            for p in self.D.parameters():
                p.requires_grad_(True)

    def step(self, frames, masked_frames, masks) -> Dict[str, torch.Tensor]:
        """Returns device scalars {rec, pool, reg, adv, dis, total}; read them outside the hot loop."""
        self.seg_a(frames, masked_frames, masks)
        d_scale = self.flat_d.all_reduce(self.pg) if self.use_gan else 1.0
        self.seg_b(d_scale)
        g_scale = self.flat_g.finish(self.pg)        # buckets were exchanged under the backward pass; this covers the rest
        self.seg_c(g_scale)
        return self._out


class GraphedStep:
    """CUDA-graph replay of a host-sync-free callable with fixed input shapes (training step or inference forward):
    ``warmup`` eager calls on a side stream, one capture, then every call = copy inputs into the static buffers +
    one graph launch (no per-kernel CPU launch cost)."""

    def __init__(self, fn, example_inputs, warmup: int = 3):
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedDPStep:
    """Data-parallel training step as THREE CUDA graphs (``GANTrainStep.seg_a/b/c``) sharing one memory pool, with the
    two NCCL all-reduces (flat D gradients, flat G gradients) issued on the same stream between the replays.  The
    collectives stay outside the captures, so nothing depends on NCCL's graph-capture support; per step the host
    issues 3 graph launches + 2 collectives instead of ~250 kernel launches."""

    def __init__(self, ts: "GANTrainStep", example_inputs, warmup: int = 3):
        self.ts = ts
        self.static_in = [torch.empty_like(t) for t in example_inputs]
        for s_, t in zip(self.static_in, example_inputs):
            s_.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                ts.step(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        world = dist.get_world_size(ts.pg) if (dist.is_available() and dist.is_initialized()) else 1
        self.scale = 1.0 / world
        self.world = world
        self.ga, self.gb, self.gc = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.ga):
            ts.seg_a(*self.static_in)
        pool = self.ga.pool()
        with torch.cuda.graph(self.gb, pool=pool):
            ts.seg_b(self.scale)
        with torch.cuda.graph(self.gc, pool=pool):
            ts.seg_c(self.scale)
        self.static_out = ts._out

    def __call__(self, *inputs):
        for s_, t in zip(self.static_in, inputs):
            if s_.data_ptr() != t.data_ptr():
                s_.copy_(t, non_blocking=True)
        ts = self.ts
        self.ga.replay()
        if self.world > 1 and ts.use_gan:
            dist.all_reduce(ts.flat_d.flat, op=dist.ReduceOp.SUM, group=ts.pg)
        self.gb.replay()
        if self.world > 1:
            dist.all_reduce(ts.flat_g.flat, op=dist.ReduceOp.SUM, group=ts.pg)
        self.gc.replay()
        return self.static_out
