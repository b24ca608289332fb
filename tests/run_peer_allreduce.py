"""torchrun worker (>= 2 GPUs on one node): p2i_peer_allreduce vs NCCL all-reduce, eager and as a CUDA-graph replay,
then one data-parallel GAN training step with the peer exchange vs the NCCL exchange.  Prints PEER_OK on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_peer_allreduce.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import synth
from p2igan_b200.peer import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def check(n):
    ar = PeerAllReduce(n)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(3):
        x = torch.randn(n, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        ar.tensor[:n].copy_(x)
        ar.all_reduce()
        torch.cuda.synchronize()
        ar.check()
        d = float((ar.tensor[:n] - ref).abs().max())
        assert d <= 1e-5 * world, (n, it, d)
        # every rank holds bitwise the same sum
        mine = ar.tensor[:n].clone()
        other = mine.clone()
        dist.broadcast(other, 0)
        assert torch.equal(mine, other), (n, it)
    # graph replay
    x = torch.randn(n, device=dev, generator=g)
    torch.cuda.synchronize()
    dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        ar.tensor[:n].copy_(x)
        ar.all_reduce()
    ref = x.clone()
    dist.all_reduce(ref)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    ar.check()
    assert float((ar.tensor[:n] - ref).abs().max()) <= 1e-5 * world
    # timing (in place, values grow: irrelevant)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ar.all_reduce()
    e1.record()
    torch.cuda.synchronize()
    t_peer = e0.elapsed_time(e1) / 20
    y = torch.randn(n, device=dev)
    for _ in range(3):
        dist.all_reduce(y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        dist.all_reduce(y)
    e1.record()
    torch.cuda.synchronize()
    t_nccl = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(f"n={n:9d} ({n * 4 / 1e6:6.1f} MB)  peer {t_peer * 1e3:7.1f} us   nccl {t_nccl * 1e3:7.1f} us", flush=True)
    return ar


keep = [check(n) for n in (1024, 1690884, 25887984)]


def check_ranges():
    """Bucketed form: disjoint float4-aligned ranges exchanged one by one (few CTAs) == one exchange of the whole buffer."""
    n = 1 << 20
    ar = PeerAllReduce(n)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    ar.tensor[:n].copy_(x)
    cuts = [0, 4096, 4096 + 36864, 500000, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ar.all_reduce_range(lo, hi - lo, 8 if hi - lo < 100000 else 32)
    torch.cuda.synchronize()
    ar.check()
    assert float((ar.tensor[:n] - ref).abs().max()) <= 1e-5 * world
    return ar


keep.append(check_ranges())


def train_losses(peer):
    from p2igan_b200 import build_discriminator, build_generator
    from p2igan_b200.train_step import GANTrainStep
    cfg = synth.make_cfg(32, 32)
    cfg["train"]["optimizer"]["lr"] = 1e-6          # two runs are compared: keep them on one trajectory (test_gpu_trainer.py _LR_NOTE)
    torch.manual_seed(2024)
    G, D = build_generator(cfg).to(dev).train(), build_discriminator(cfg).to(dev).train()
    ts = GANTrainStep(cfg, G, D, peer_exchange=peer)
    out = []
    for i in range(2):
        b = tuple(t.to(dev) for t in synth.make_batch(2, 16, 32, 32, 12, 500 + 10 * rank + i))
        out.append({k: float(v) for k, v in ts.step(*b).items()})
    fp = torch.cat([p.detach().reshape(-1) for p in G.parameters()])
    return out, fp


ln, pn = train_losses(False)
lp, pp = train_losses(True)
for a, b in zip(ln, lp):
    for k in a:
        assert abs(a[k] - b[k]) < 2e-3 * abs(a[k]) + 1e-5, (k, a[k], b[k])
d = (pn - pp).abs()
assert float(d.max()) <= 3e-5 and float(d.mean()) < 2e-7, (float(d.max()), float(d.mean()))
# data-parallel invariant: parameters stay identical across ranks
ref = pp.clone()
dist.broadcast(ref, 0)
assert torch.equal(ref, pp), "parameters diverged across ranks"


def dp_equivalence():
    """N ranks x b events == 1 rank x N*b events: the REAL GANTrainStep at 32x32 with the bucketed, overlapped
    peer exchange (eager and as ONE CUDA graph) against the same step on the whole batch without any exchange.  Compared:
    the exchanged flat gradients (x 1/world) of G and D and the mean of the ranks' losses, to fp32 summation-order tolerance
    (activations are per-event, so only the order of the batch reductions differs)."""
    from p2igan_b200 import build_discriminator, build_generator
    from p2igan_b200.train_step import GANTrainStep, GraphedStep
    cfg = synth.make_cfg(32, 32)
    cfg["train"]["optimizer"]["lr"] = 1e-6
    b = 2
    full = tuple(t.to(dev) for t in synth.make_batch(b * world, 16, 32, 32, 12, 900))
    mine = tuple(t[rank * b:(rank + 1) * b].contiguous() for t in full)
    solo = None
    for r in range(world):                      # every rank takes part in creating every one-rank group
        g_ = dist.new_group([r])
        if r == rank:
            solo = g_

    def run(batch, **kw):
        torch.manual_seed(2024)
        G, D = build_generator(cfg).to(dev).train(), build_discriminator(cfg).to(dev).train()
        ts = GANTrainStep(cfg, G, D, **kw)
        out = {k: float(v) for k, v in ts.step(*batch).items()}
        torch.cuda.synchronize()
        return ts, out, ts.flat_g.flat[:ts.flat_g.n].clone(), ts.flat_d.flat[:ts.flat_d.n].clone()

    _, l1, g1, d1 = run(full, process_group=solo)
    os.environ["P2I_BUCKETED"] = "1"            # the bucketed form is opt-in for N > 1 (train_step.py); exercise it here
    ts, lp, gp, dp = run(mine, peer_exchange=True)
    assert ts.peer_exchange and getattr(ts.G, "_bucket_hook", None) is not None
    for name, a_, b_ in (("G", gp / world, g1), ("D", dp / world, d1)):
        rel = float((a_ - b_).norm() / b_.norm())
        # activations are per event, but run-to-run noise of the step itself is ~2e-3 (fp32 atomics reorder the power iteration
        # and the weight gradients; a last-bit change of sigma flips bf16 roundings).  A lost bucket or a wrong 1/world would be O(1).
        assert rel < 1e-2, (name, rel)
    for k in ("rec", "pool", "dis"):
        t = torch.tensor([lp[k]], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(float(t) / world - l1[k]) < 1e-4 * abs(l1[k]) + 1e-6, (k, float(t) / world, l1[k])
    # the same data-parallel step captured as ONE CUDA graph (bucket exchanges on the comm stream inside the capture)
    gs = GraphedStep(lambda a, b_, c: ts.step(a, b_, c)["total"], mine, warmup=1)
    for _ in range(2):
        gs(*mine)
    torch.cuda.synchronize()
    ts.flat_g.peer.check(); ts.flat_d.peer.check()
    fp = torch.cat([p.detach().reshape(-1) for p in ts.G.parameters()])
    ref_ = fp.clone()
    dist.broadcast(ref_, 0)
    assert torch.equal(ref_, fp), "parameters diverged across ranks after graph replays"
    os.environ["P2I_BUCKETED"] = "0"
    ts0, l0, g0, d0 = run(mine, peer_exchange=True)          # default form: one whole-buffer exchange after the backward pass
    assert getattr(ts0.G, "_bucket_hook", None) is None
    rel = float((g0 / world - g1).norm() / g1.norm())
    assert rel < 1e-2, ("G unbucketed", rel)


dp_equivalence()
dist.barrier()
if rank == 0:
    print("PEER_OK", flush=True)
dist.destroy_process_group()
