import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch, synth
from p2igan_b200 import build_generator, build_discriminator
from p2igan_b200.train_step import GANTrainStep
dev = "cuda:0"
cfg = synth.make_cfg(128, 128)
torch.manual_seed(2024)
G = build_generator(cfg).to(dev).train(); D = build_discriminator(cfg).to(dev).train()
ts = GANTrainStep(cfg, G, D)
B = 16
batch = tuple(t.to(dev) for t in synth.make_batch(B, 16, 128, 128, 79, 1))
for i in range(3): ts.step(*batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5): ts.step(*batch)
t1 = time.perf_counter()   # CPU issue time (async)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"cpu issue {1e3*(t1-t0)/5:.2f} ms/step, total {1e3*(t2-t0)/5:.2f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): ts.step(*batch)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(ka, key=lambda e: -e.device_time_total)[:45]
tot = sum(e.device_time_total for e in ka if e.device_type.name == "CUDA" or e.device_time_total > 0)
for e in rows:
    if e.device_time_total > 0:
        print(f"{e.key[:70]:70s} n={e.count:5d} cuda={e.device_time_total/3e3:8.3f} ms/step")
print("sum cuda (ms/step):", sum(e.self_device_time_total for e in ka)/3e3)
