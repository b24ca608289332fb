"""Per-role wait cycles of conv_halo_kernel (diagnostics build -DHALO_PROF, block 0 prints from the device):
    P2I_HALO_PROF=1 python -c "from p2igan_b200.build import build; build()"     # -> libp2i_sm100a_prof.so
    P2I_LIB_PATH=.../libp2i_sm100a_prof.so python tools/halo_prof.py
Each case is launched 3 times (warm L2, as inside the step); the device lines of every launch follow its header."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch

from p2igan_b200 import ops

dev = "cuda:0"
B = 16
g = torch.Generator(device=dev).manual_seed(1)
levels = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
print("P2I_HALO_DBG =", os.environ.get("P2I_HALO_DBG", "0"), flush=True)
for lvl, C in enumerate((64, 128, 256, 512)):
    if lvl not in levels:
        continue
    hw = 128 >> lvl
    x = torch.randn(B, hw, hw, C, device=dev, generator=g).to(torch.bfloat16)
    r = torch.randn(B, hw, hw, C, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(9, C, C, device=dev, generator=g) / (9 * C) ** 0.5).to(torch.bfloat16)
    for what, kw in (("relu", dict(relu=True)), ("+res", dict(residual=r))):
        for it in range(3):
            torch.cuda.synchronize()
            print(f"=== L{lvl} {C}ch {hw}x{hw} {what} launch {it}", flush=True)
            ops.conv2d_cl(x, w, **kw)
            torch.cuda.synchronize()
