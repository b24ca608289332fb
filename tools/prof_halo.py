"""Run single halo-conv launches of a -DHALO_PROF build (P2I_HALO_PROF=1 python .../build.py -f) and let block 0
print its per-role wait cycles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch
from p2igan_b200._lib import LIB
from p2igan_b200.disc_ops import conv_desc, conv_igemm
dev = "cuda:0"; bf = torch.bfloat16
B = 16
LIB.call("p2i_set_conv_impl", int(sys.argv[1]) if len(sys.argv) > 1 else 4)
for lvl, C in enumerate((64, 128, 256, 512)):
    hw = 128 >> lvl
    x = torch.randn(B, 1, hw, hw, C, device=dev).to(bf)
    w = (torch.randn(9, C, C, device=dev) / (9 * C) ** 0.5).to(bf)
    r = torch.randn(B, hw, hw, C, device=dev).to(bf)
    y = torch.empty(B, hw, hw, C, device=dev, dtype=bf)
    for aux in (None, r):
        desc = conv_desc(B, 1, 1, hw, hw, C, C, 1, 3, 1, 0, act=0 if aux is not None else 1)
        torch.cuda.synchronize()
        print(f"--- L{lvl} C={C} {hw}x{hw} {'residual' if aux is not None else 'relu'}", flush=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        conv_igemm(x, w, desc, residual=aux, out=y)
        e1.record()
        torch.cuda.synchronize()
        print(f"    kernel {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
