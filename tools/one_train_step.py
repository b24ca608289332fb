"""Run a few eager training steps (B=16, 16x128x128); the last one between cudaProfilerStart/Stop for
`ncu --profile-from-start off` launch lists."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch
import synth
from p2igan_b200 import build_discriminator, build_generator
from p2igan_b200.train_step import GANTrainStep

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = synth.make_cfg(128, 128)
torch.manual_seed(2024)
G = build_generator(cfg).to(dev).train()
D = build_discriminator(cfg).to(dev).train()
ts = GANTrainStep(cfg, G, D)
batch = tuple(t.to(dev) for t in synth.make_batch(B, 16, 128, 128, 79, 1))
for _ in range(2):
    ts.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off (the backward runs on autograd's thread)
o = ts.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: float(v) for k, v in o.items()})
