"""A/B of the two weight-gradient kernels (first generation: scalar REDs, lanes = ci; second generation: flipped GEMM
with 16-byte vector REDs and the all-taps mode for 64->64 layers): max-abs difference and CUDA-event time per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch

from p2igan_b200._lib import LIB
from p2igan_b200.disc_ops import conv_desc, conv_wgrad

dev = "cuda:0"
bf = torch.bfloat16


def case(name, samples, T_in, T_out, H, W, Cin, Cout, kt, k, pad, pad_t, stride_t=1, iters=20):
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.randn(samples, T_in, H, W, Cin, device=dev, generator=g).to(bf) for _ in range(3)]
    dys = [torch.randn(samples, T_out, H, W, Cout, device=dev, generator=g).to(bf) for _ in range(3)]
    desc = conv_desc(samples, T_in, T_out, H, W, Cin, Cout, kt, k, pad, pad_t, stride_t)
    res = {}
    for impl in (1, IMPL2):
        LIB.call("p2i_set_wgrad_impl", impl)
        dW = torch.zeros(kt * k * k, Cout, Cin, device=dev)
        conv_wgrad(xs[0], dys[0], dW, desc)
        torch.cuda.synchronize()
        ref = dW.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            conv_wgrad(xs[i % 3], dys[i % 3], dW, desc)
        e1.record()
        torch.cuda.synchronize()
        res[impl] = (ref, e0.elapsed_time(e1) / iters * 1e3)
    LIB.call("p2i_set_wgrad_impl", 0)
    fl = 2.0 * samples * T_out * H * W * Cout * Cin * kt * k * k
    a, ta = res[1]
    b, tb = res[IMPL2]
    scale = float(a.abs().max())
    print(f"{name:36s} | gen1 {ta:6.1f} us {fl / ta / 1e6:6.0f} TF/s | gen2 {tb:6.1f} us {fl / tb / 1e6:6.0f} TF/s | "
          f"max|d| {float((a - b).abs().max()):.2e} (scale {scale:.1e})", flush=True)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
IMPL2 = int(sys.argv[2]) if len(sys.argv) > 2 else 2      # second column: 2 = experimental second-generation kernels
print(f"B = {B}")
for lvl, C in enumerate((64, 128, 256, 512)):
    hw = 128 >> lvl
    case(f"G L{lvl} {C}ch {hw}x{hw} k3", B, 1, 1, hw, hw, C, C, 1, 3, 1, 0)
for lvl, C in ((2, 256), (3, 512)):
    hw = 128 >> lvl
    case(f"UP proj {C}->{C // 2} {hw}x{hw} 1x1", B, 1, 1, hw, hw, C, C // 2, 1, 1, 0, 0)
case("d2d.2 256->128 64 k2", B, 1, 1, 64, 64, 256, 128, 1, 2, 1, 0)
case("d2d.4 512->256 32 k2", B, 1, 1, 32, 32, 512, 256, 1, 2, 1, 0)
case("d2d.6 256->256 32 k3", B, 1, 1, 32, 32, 256, 256, 1, 3, 1, 0)
case("d3d.4 256->128 T16 16 kt3k2", B, 16, 16, 16, 16, 256, 128, 3, 2, 1, 1)
case("d3d.6 128->128 T16->8 16 kt3k3 st2", B, 16, 8, 16, 16, 128, 128, 3, 3, 1, 1, stride_t=2)
case("ragged 64ch 24x40", 1, 1, 1, 24, 40, 64, 64, 1, 3, 1, 0, iters=3)
case("small 64ch 8x8", 2, 1, 1, 8, 8, 64, 64, 1, 3, 1, 0, iters=3)
case("ragged 128ch 20x12", 2, 1, 1, 20, 12, 128, 128, 1, 3, 1, 0, iters=3)
