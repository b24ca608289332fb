"""Thin tensor-level wrappers over the C ABI (one function per entry point of include/p2i_b200.h).

Tensors are PyTorch-owned device memory; PyTorch supplies the allocator and the current stream only.
"cl" = channels-last activations, shape [B, H, W, C], bf16, contiguous.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import LIB, ptr, require_cuda, stream

Tensor = torch.Tensor


def _chk(t: Tensor, dtype, name: str) -> Tensor:
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


# ------------------------------------------------------------------------------------------- InputBlock
def points_extract(masks: Tensor, cap: Optional[int] = None):
    """masks [B,T,H,W] f32 -> (pts [B,cap] i32, counts [B] i32, src [B] i32)."""
    require_cuda(masks)
    B, T, H, W = masks.shape
    cap = int(cap or T * H * W)
    pts = torch.empty(B, cap, dtype=torch.int32, device=masks.device)
    counts = torch.empty(B, dtype=torch.int32, device=masks.device)
    src = torch.empty(B, dtype=torch.int32, device=masks.device)
    LIB.call("p2i_points_extract", ptr(_chk(masks, torch.float32, "masks")), B, T, H, W, ptr(pts), ptr(counts), ptr(src),
             cap, stream())
    return pts, counts, src


def gate_points_fwd(masked: Tensor, pts: Tensor, counts: Tensor, w0, b0, w1, b1, save_l1: bool = False):
    B, T, H, W = masked.shape
    cap = pts.shape[1]
    vals = torch.zeros(B, cap, dtype=torch.float32, device=masked.device)
    l1 = torch.empty(B, cap, 16, dtype=torch.float32, device=masked.device) if save_l1 else None
    LIB.call("p2i_gate_points_fwd", ptr(_chk(masked, torch.float32, "masked")), ptr(pts), ptr(counts), cap,
             ptr(_chk(w0, torch.float32, "w0")), ptr(b0), ptr(_chk(w1, torch.float32, "w1")), ptr(b1), ptr(vals), ptr(l1),
             B, T, H, W, stream())
    return vals, l1


def gate_points_bwd(masked, pts, counts, w0, b0, w1, b1, dvals, dw0, db0, dw1, db1):
    """Accumulates (atomics) into dw0 [16,16,1], db0 [16], dw1, db1 (float32, contiguous)."""
    B, T, H, W = masked.shape
    cap = pts.shape[1]
    LIB.call("p2i_gate_points_bwd", ptr(masked), ptr(pts), ptr(counts), cap, ptr(w0), ptr(b0), ptr(w1), ptr(b1),
             ptr(_chk(dvals, torch.float32, "dvals")), ptr(dw0), ptr(db0), ptr(dw1), ptr(db1), B, T, H, W, stream())


class IdwTableCache:
    """Device-resident copy of sample 0's neighbour table, reused while its point pattern is unchanged (decided on
    the device by p2i_idw_cache_check: no host synchronisation, CUDA-graph capturable)."""

    def __init__(self, shape: Tuple[int, int, int], tau: float, cap: int, device):
        T, H, W = shape
        self.key = (T, H, W, float(tau), cap, str(device))
        self.pts = torch.zeros(cap, dtype=torch.int32, device=device)
        self.count = torch.full((1,), -1, dtype=torch.int32, device=device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.idx = torch.zeros(T * H * W, 4, dtype=torch.int32, device=device)
        self.w = torch.zeros(T * H * W, 4, dtype=torch.float32, device=device)

    def check(self, pts: Tensor, counts: Tensor):
        LIB.call("p2i_idw_cache_check", ptr(pts), ptr(counts), pts.shape[1], ptr(self.pts), ptr(self.count), ptr(self.flag),
                 stream())


def idw_knn_fwd(pts, vals, counts, src, shape: Tuple[int, int, int], tau: float, table=None,
                cache: Optional[IdwTableCache] = None):
    """-> (out [B,T,H,W] f32, (nbr_idx, nbr_w)).  Pass `table` to reuse a neighbour table (no search); pass `cache`
    (after cache.check) to let sample 0's rows come from / go to the cross-call cache."""
    T, H, W = shape
    B, cap = pts.shape
    Q = T * H * W
    out = torch.empty(B, T, H, W, dtype=torch.float32, device=pts.device)
    search = table is None
    if search:
        table = (torch.empty(B, Q, 4, dtype=torch.int32, device=pts.device),
                 torch.empty(B, Q, 4, dtype=torch.float32, device=pts.device))
    LIB.call("p2i_idw_knn_fwd", ptr(pts), ptr(vals), ptr(counts), ptr(src), cap, ptr(out), ptr(table[0]), ptr(table[1]),
             B, T, H, W, float(tau), 1 if search else 0, ptr(cache.flag if cache else None),
             ptr(cache.idx if cache else None), ptr(cache.w if cache else None), stream())
    return out, table


def idw_knn_bwd(dout, table, counts, src, cap: int):
    B, T, H, W = dout.shape
    dvals = torch.zeros(B, cap, dtype=torch.float32, device=dout.device)
    LIB.call("p2i_idw_knn_bwd", ptr(_chk(dout, torch.float32, "dout")), ptr(table[0]), ptr(table[1]), ptr(counts), ptr(src),
             ptr(dvals), cap, B, T, H, W, stream())
    return dvals


# ------------------------------------------------------------------------------------------- convolutions
def conv2d_cl(x: Tensor, w: Tensor, residual: Optional[Tensor] = None, relu: bool = False,
              out: Optional[Tensor] = None, direct: bool = False, mask: Optional[Tensor] = None,
              bias: Optional[Tensor] = None, leaky: bool = False) -> Tensor:
    """x [B,H,W,Cin] bf16, w [k*k,Cout,Cin] bf16 -> [B,H,W,Cout] bf16 (tcgen05 implicit GEMM)."""
    require_cuda(x, w)
    B, H, W, Cin = x.shape
    taps, Cout, Cin2 = w.shape
    if Cin2 != Cin or taps not in (1, 9):
        raise ValueError(f"conv2d_cl: weight {tuple(w.shape)} does not match input channels {Cin}")
    k = 3 if taps == 9 else 1
    if out is None:
        out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=x.device)
    LIB.call("p2i_conv2d_direct_fwd" if direct else "p2i_conv2d_igemm_fwd", ptr(_chk(x, torch.bfloat16, "x")),
             ptr(_chk(w, torch.bfloat16, "w")), ptr(residual), ptr(mask), ptr(bias), ptr(out), B, H, W, Cin, Cout, k,
             (1 if relu else 0) | (2 if leaky else 0), stream())
    return out


def conv2d_wgrad(x: Tensor, dy: Tensor, ksize: int, out: Optional[Tensor] = None) -> Tensor:
    """x [B,H,W,Cin] bf16, dy [B,H,W,Cout] bf16 -> dW f32 [k*k,Cout,Cin] (accumulated into `out` if given)."""
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    if out is None:
        out = torch.zeros(ksize * ksize, Cout, Cin, dtype=torch.float32, device=x.device)
    LIB.call("p2i_conv2d_wgrad", ptr(_chk(x, torch.bfloat16, "x")), ptr(_chk(dy, torch.bfloat16, "dy")), ptr(out), B, H, W,
             Cin, Cout, ksize, stream())
    return out


def doconv_compose(table_dev: Tensor, n_layers: int, max_channels: int) -> None:
    LIB.call("p2i_doconv_compose_fwd", ptr(table_dev), n_layers, max_channels, stream())


def doconv_compose_stem(W: Tensor, D: Tensor, D_diag: Tensor) -> Tensor:
    out = torch.empty(64, 4, 9, dtype=torch.float32, device=W.device)
    LIB.call("p2i_doconv_compose_stem_fwd", ptr(_chk(W, torch.float32, "W")), ptr(_chk(D, torch.float32, "D")),
             ptr(_chk(D_diag, torch.float32, "D_diag")), ptr(out), stream())
    return out


# ------------------------------------------------------------------------------------------- generator glue
def stem_fwd(x: Tensor, w: Tensor) -> Tensor:
    B, C, H, W = x.shape
    assert C == 16
    y = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=x.device)
    LIB.call("p2i_stem_fwd", ptr(_chk(x, torch.float32, "x")), ptr(_chk(w, torch.float32, "w")), ptr(y), B, H, W, stream())
    return y


def pyramid_fwd(stem: Tensor):
    B, H, W, C = stem.shape
    assert C == 64
    x4 = torch.empty(B, H // 4, W // 4, 256, dtype=torch.bfloat16, device=stem.device)
    x8 = torch.empty(B, H // 8, W // 8, 512, dtype=torch.bfloat16, device=stem.device)
    LIB.call("p2i_pyramid_fwd", ptr(_chk(stem, torch.bfloat16, "stem")), ptr(x4), ptr(x8), B, H, W, stream())
    return x4, x8


def downsample_dup_fwd(x: Tensor) -> Tensor:
    """x [B,C,H,W] f32 -> [B,2C,H/2,W/2] f32: max_pool2d(2,2) + channel duplication (layer.py:205-214)."""
    require_cuda(x)
    B, C, H, W = x.shape
    y = torch.empty(B, 2 * C, H // 2, W // 2, dtype=torch.float32, device=x.device)
    LIB.call("p2i_downsample_dup_fwd", ptr(_chk(x, torch.float32, "x")), ptr(y), B, C, H, W, stream())
    return y


def downsample_dup_bwd(x: Tensor, dy: Tensor) -> Tensor:
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    LIB.call("p2i_downsample_dup_bwd", ptr(_chk(x, torch.float32, "x")), ptr(_chk(dy, torch.float32, "dy")), ptr(dx), B, C, H, W,
             stream())
    return dx


def upmod_fwd(z: Tensor, pos: Tensor, bias: Tensor, skip: Optional[Tensor] = None) -> Tensor:
    B, h, w, C = z.shape
    out = torch.empty(B, 2 * h, 2 * w, C, dtype=torch.bfloat16, device=z.device)
    LIB.call("p2i_upmod_fwd", ptr(_chk(z, torch.bfloat16, "z")), ptr(_chk(pos, torch.float32, "pos")),
             ptr(_chk(bias, torch.float32, "bias")), ptr(skip), ptr(out), B, h, w, C, stream())
    return out


def head_fwd(x: Tensor, w: Tensor, want_pre: bool = False):
    B, H, W, C = x.shape
    assert C == 64
    out = torch.empty(B, 16, H, W, dtype=torch.float32, device=x.device)
    pre = torch.empty_like(out) if want_pre else None
    LIB.call("p2i_head_fwd", ptr(_chk(x, torch.bfloat16, "x")), ptr(_chk(w, torch.float32, "w")), ptr(out), ptr(pre), B, H, W,
             stream())
    return (out, pre) if want_pre else out


def to_cl(x: Tensor) -> Tensor:
    """NCHW f32 -> NHWC bf16."""
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x.device)
    LIB.call("p2i_nchw_f32_to_nhwc_bf16", ptr(_chk(x, torch.float32, "x")), ptr(y), B, C, H, W, stream())
    return y


def from_cl(x: Tensor) -> Tensor:
    """NHWC bf16 -> NCHW f32."""
    B, H, W, C = x.shape
    y = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    LIB.call("p2i_nhwc_bf16_to_nchw_f32", ptr(_chk(x, torch.bfloat16, "x")), ptr(y), B, C, H, W, stream())
    return y


# ------------------------------------------------------------------------------------------- backward glue
def head_bwd(dout: Tensor, out: Tensor, x: Tensor, w: Tensor):
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    dw = torch.zeros(16, 16, dtype=torch.float32, device=x.device)
    LIB.call("p2i_head_bwd", ptr(_chk(dout, torch.float32, "dout")), ptr(_chk(out, torch.float32, "out")), ptr(x), ptr(w),
             ptr(dx), ptr(dw), B, H, W, stream())
    return dx, dw


def upmod_bwd(z: Tensor, pos: Tensor, bias: Tensor, dout: Tensor, dbias: Tensor, dpos: Tensor):
    """-> dz [B,h,w,C] bf16; accumulates into dbias f32 [C] and dpos f32 [...,2h,2w]."""
    B, h, w, C = z.shape
    scratch = torch.empty_like(dout)
    dz = torch.empty_like(z)
    LIB.call("p2i_upmod_bwd", ptr(z), ptr(_chk(pos, torch.float32, "pos")), ptr(_chk(bias, torch.float32, "bias")),
             ptr(_chk(dout, torch.bfloat16, "dout")), ptr(scratch), ptr(dz), ptr(_chk(dbias, torch.float32, "dbias")),
             ptr(_chk(dpos, torch.float32, "dpos")), B, h, w, C, stream())
    return dz


def pyramid_bwd(stem: Tensor, dx4: Tensor, dx8: Tensor) -> Tensor:
    B, H, W, C = stem.shape
    d = torch.empty_like(stem)
    LIB.call("p2i_pyramid_bwd", ptr(stem), ptr(_chk(dx4, torch.bfloat16, "dx4")), ptr(_chk(dx8, torch.bfloat16, "dx8")), ptr(d),
             B, H, W, stream())
    return d


def stem_bwd(dy: Tensor, x: Tensor, w: Tensor):
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    dw = torch.zeros(64, 4, 9, dtype=torch.float32, device=x.device)
    LIB.call("p2i_stem_bwd", ptr(_chk(dy, torch.bfloat16, "dy")), ptr(x), ptr(w), ptr(dx), ptr(dw), B, H, W, stream())
    return dx, dw


def stem_bwd_dx(dy: Tensor, x: Tensor, w: Tensor) -> Tensor:
    """Input gradient of the stem only (the weight gradient is an independent kernel: stem_bwd_dw)."""
    B, C, H, W = x.shape
    dx = torch.empty_like(x)
    LIB.call("p2i_stem_bwd", ptr(_chk(dy, torch.bfloat16, "dy")), ptr(x), ptr(w), ptr(dx), None, B, H, W, stream())
    return dx


def stem_bwd_dw(dy: Tensor, x: Tensor, w: Tensor, dw: Tensor) -> Tensor:
    """Weight gradient of the stem, accumulated into the zero-filled dw f32 [64,4,9]."""
    B, C, H, W = x.shape
    LIB.call("p2i_stem_bwd", ptr(_chk(dy, torch.bfloat16, "dy")), ptr(x), ptr(w), None, ptr(_chk(dw, torch.float32, "dw")), B, H, W,
             stream())
    return dw


def doconv_compose_bwd(table_dev: Tensor, n_layers: int, max_channels: int) -> None:
    LIB.call("p2i_doconv_compose_bwd", ptr(table_dev), n_layers, max_channels, stream())


def doconv_compose_stem_bwd(W, D, D_diag, dDoW, dW, dD):
    """Accumulates into dW [64,4,9] and dD [16,9,9]."""
    LIB.call("p2i_doconv_compose_stem_bwd", ptr(W), ptr(D), ptr(D_diag), ptr(_chk(dDoW, torch.float32, "dDoW")), ptr(dW),
             ptr(dD), stream())
