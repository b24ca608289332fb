"""A/B of the two implicit-GEMM conv kernels (legacy per-kx boxes vs halo box) on the step's layer shapes:
max-abs / rel-L2 difference of the outputs and CUDA-event time per launch (inputs rotate over 3 buffers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch

from p2igan_b200._lib import LIB
from p2igan_b200.disc_ops import conv_desc, conv_igemm

dev = "cuda:0"
bf = torch.bfloat16


def out_shape(F, H, W, Cout, out_mode):
    if out_mode == 0:
        return (F, H, W, Cout)
    if out_mode == 1:
        return (F, H // 2, W // 2, 4 * Cout)
    return (F, 2 * H, 2 * W, Cout // 4)


def case(name, samples, T_in, T_out, H, W, Cin, Cout, kt, k, pad, pad_t, stride_t=1, tt=0, act=0, aux=None, out_mode=0,
         bias=False, iters=20):
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.randn(samples, T_in, H, W, Cin, device=dev, generator=g).to(bf) for _ in range(3)]
    w = (torch.randn(kt * k * k, Cout, Cin, device=dev, generator=g) / (Cin * k * k * kt) ** 0.5).to(bf)
    F = samples * T_out
    auxt = torch.randn(F, H, W, Cout, device=dev, generator=g).to(bf) if aux else None
    b = torch.randn(Cout, device=dev, generator=g) if bias else None
    mask_mode = {None: 0, "res": 0, "mask": 1, "lmask": 2}[aux]
    desc = conv_desc(samples, T_in, T_out, H, W, Cin, Cout, kt, k, pad, pad_t, stride_t, tt, act, mask_mode, out_mode)
    res = {}
    for impl in IMPLS:
        LIB.call("p2i_set_conv_impl", impl)
        outs = [torch.zeros(out_shape(F, H, W, Cout, out_mode), device=dev, dtype=bf) for _ in range(3)]

        def run(i):
            conv_igemm(xs[i % 3], w, desc, residual=auxt if aux == "res" else None,
                       mask=auxt if aux in ("mask", "lmask") else None, bias=b, out=outs[i % 3])
        try:
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
        except RuntimeError as e:
            res[impl] = (None, str(e)[:80])
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        res[impl] = (outs[0].float(), e0.elapsed_time(e1) / iters * 1e3)
    LIB.call("p2i_set_conv_impl", 0)
    fl = 2.0 * F * H * W * Cout * Cin * kt * k * k
    line = f"{name:34s}"
    for impl in IMPLS:
        o, t = res[impl]
        line += f" | {NAMES[impl]}: " + (f"{t:6.1f} us {fl / t / 1e6:6.0f} TF/s" if o is not None else f"n/a ({t[:40]})")
        if impl != IMPLS[0] and o is not None and res[IMPLS[0]][0] is not None:
            a = res[IMPLS[0]][0]
            line += f" d={float((a - o).abs().max()):.1e}"
    print(line, flush=True)


IMPLS = (1, 3, 4)
NAMES = {1: "legacy", 2: "halo", 3: "halo1", 4: "halo2"}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
print(f"B = {B}")
for lvl, C in enumerate((64, 128, 256, 512)):
    hw = 128 >> lvl
    case(f"G L{lvl} {C}ch {hw}x{hw} relu", B, 1, 1, hw, hw, C, C, 1, 3, 1, 0, act=1)
    case(f"G L{lvl} {C}ch {hw}x{hw} +res", B, 1, 1, hw, hw, C, C, 1, 3, 1, 0, aux="res")
    case(f"G L{lvl} {C}ch {hw}x{hw} mask (dgrad)", B, 1, 1, hw, hw, C, C, 1, 3, 1, 0, aux="mask")
for lvl, C in ((1, 128), (2, 256), (3, 512)):
    hw = 128 >> lvl
    case(f"UP proj {C}->{C // 2} {hw}x{hw} 1x1", B, 1, 1, hw, hw, C, C // 2, 1, 1, 0, 0)
case("d2d.0 64->64 128 k3 s2d-pack", B, 1, 1, 128, 128, 64, 64, 1, 3, 1, 0, act=2, out_mode=1, bias=True)
case("d2d.2 256->128 64 k2 s2d-pack", B, 1, 1, 64, 64, 256, 128, 1, 2, 1, 0, act=2, out_mode=1, bias=True)
case("d2d.4 512->256 32 k2", B, 1, 1, 32, 32, 512, 256, 1, 2, 1, 0, act=2, bias=True)
case("d2d.6 256->256 32 k3", B, 1, 1, 32, 32, 256, 256, 1, 3, 1, 0, act=2, bias=True)
case("d3d.2 128->64 T16 32 kt3k2 pack", B, 16, 16, 32, 32, 128, 64, 3, 2, 1, 1, act=2, out_mode=1, bias=True)
case("d3d.4 256->128 T16 16 kt3k2", B, 16, 16, 16, 16, 256, 128, 3, 2, 1, 1, act=2, bias=True)
case("d3d.6 128->128 T16->8 16 kt3k3 st2", B, 16, 8, 16, 16, 128, 128, 3, 3, 1, 1, stride_t=2, act=2, bias=True)
# data-gradient forms (transposed weights: Cin/Cout swapped, k=2 pads on the other side, unpack / lmask)
case("d2d.2 dgrad 128->256 64 k2 unpack", B, 1, 1, 64, 64, 128, 256, 1, 2, 0, 0, aux="lmask", out_mode=2)
case("d3d.2 dgrad 64->128 T16 32 unpack", B, 16, 16, 32, 32, 64, 128, 3, 2, 0, 1, aux="lmask", out_mode=2)
case("d3d.6 dgrad 128->128 T8->16 tmode", B, 8, 16, 16, 16, 128, 128, 3, 3, 1, 1, stride_t=2, tt=1, aux="lmask")
# ragged / small shapes (parity only)
case("ragged 64ch 24x40", 1, 1, 1, 24, 40, 64, 64, 1, 3, 1, 0, act=1, iters=3)
case("small 512ch 8x8 +res", 2, 1, 1, 8, 8, 512, 512, 1, 3, 1, 0, aux="res", iters=3)
