"""Kernel-time vs wall-time of ONE CUDA-graph replay of the training step (B=16): how much of the step is launch gaps /
tails between dependent kernels.  Uses the torch profiler (CUPTI) on graph replays."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch
import synth
from p2igan_b200 import build_discriminator, build_generator
from p2igan_b200.train_step import GANTrainStep, GraphedStep
from torch.profiler import ProfilerActivity, profile

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:                      # data-parallel step under torchrun: peer-memory exchange, rank 0 reports
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
cfg = synth.make_cfg(128, 128)
torch.manual_seed(2024)
G = build_generator(cfg).to(dev).train()
D = build_discriminator(cfg).to(dev).train()
ts = GANTrainStep(cfg, G, D, peer_exchange=world > 1)
batch = tuple(t.to(dev) for t in synth.make_batch(16, 16, 128, 128, 79, 1 + rank))
gs = GraphedStep(lambda a, b, c: ts.step(a, b, c)["total"], batch, warmup=3)
for _ in range(3):
    gs(*batch)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs(*batch)
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
wall = ev[-1].time_range.end - ev[0].time_range.start
gaps = sorted(((ev[i + 1].time_range.start - ev[i].time_range.end, ev[i].name[:40], ev[i + 1].name[:40]) for i in range(len(ev) - 1)), reverse=True)
print(f"kernels {len(ev)}  busy {busy / 1e3:.3f} ms  wall {wall / 1e3:.3f} ms  gaps {(wall - busy) / 1e3:.3f} ms  mean gap {(wall - busy) / max(1, len(ev) - 1):.2f} us")
for g in gaps[:12]:
    print(f"  gap {g[0]:7.2f} us  after {g[1]}  before {g[2]}")
agg = {}
for e in ev:
    k = e.name.split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{v[1]:9.1f} us  n={v[0]:3d}  {k}")
# full timeline (start relative to the first kernel, duration, stream): where the critical path and the overlaps are
if len(sys.argv) > 1:
    t0 = ev[0].time_range.start
    with open(sys.argv[1], "w") as f:
        for e in ev:
            f.write(f"{(e.time_range.start - t0):10.2f} {(e.time_range.end - e.time_range.start):8.2f} s{getattr(e, 'device_resource_id', -1)} {e.name.split('(')[0][:70]}\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
