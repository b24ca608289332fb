"""CPU oracle for the P2I-GAN hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 (and, for the IDW tie analysis, fp64/integer) PyTorch-on-CPU
restatement of the arithmetic of the reference's hot path.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``p2i-gan-benchmark_b200/`` imports it; the product path
fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, executed in the build container by
``tests/golden/make_golden.py`` (imports /root/reference with an in-memory torchmetrics stub) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function below
against those vectors; ``batch_post_process`` is pinned by ``tests/golden/make_golden_data.py`` (the reference's own
``Dataset.post_process``) and ``tests/test_batch_prep_oracle.py``.
PARITY UNPINNED for exactly one function: ``ssim_per_image`` restates torchmetrics 1.0.3 (a third-party dependency of
the reference, pinned in its uv.lock:1162-1164, absent from /root/reference and from this image) from its published
algorithm; the reference holds no SSIM value to check it against.

Everything is functional: parameters come in as a ``state_dict``-style mapping with the
reference's key names, so the same weights can be fed to the CUDA modules and to the oracle.

Reference citations (file:line under /root/reference):
  gate                      p2igan_bench/modules/layer.py:296-304
  point extraction          p2igan_bench/modules/layer.py:325-344
  idw                       p2igan_bench/modules/layer.py:246-293
  DO-Conv weight compose    p2igan_bench/modules/deconv_pytorch.py:111-132
  ResBlock / EBlock         p2igan_bench/modules/layer.py:126-135, models/p2igan.py:176-183
  pyramid                   p2igan_bench/modules/layer.py:200-214
  UPPos                     p2igan_bench/modules/layer.py:384-399
  generator forward         p2igan_bench/models/p2igan.py:72-112
  discriminator forward     p2igan_bench/models/p2igan.py:157-173
  spectral norm             torch.nn.utils.spectral_norm (call sites layer.py:402-407, p2igan.py:141)
  reconstruction loss       p2igan_bench/modules/losses.py:38-85
  adversarial loss          p2igan_bench/modules/losses.py:192-253
  metrics                   p2igan_bench/metrics/metric.py:16-183
  train step order          scripts/train.py:240-326
  sliding-window inference  scripts/infer.py:217-245
  batch preparation         p2igan_bench/data/sti_dataset.py:203-243, scripts/train.py:468-473
  SSIM                      p2igan_bench/metrics/metric.py:36,55-56,69 -> torchmetrics (unpinned, see above)
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Mapping[str, Tensor]

# --------------------------------------------------------------------------------------
# InputBlock: gate, point extraction, inverse-distance kNN
# --------------------------------------------------------------------------------------


def gate(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """One AttentionBlock on rows of length 16: relu(x + x * (W x + b)).  x: [P, 16]."""
    g = x @ weight.reshape(weight.shape[0], weight.shape[1]).t() + bias
    return torch.relu(x + x * g)


def gated_frames(sd: State, masked: Tensor, prefix: str = "input.layers") -> Tensor:
    """Both gates applied at every pixel.  masked: [B,16,H,W] -> [B,16,H,W]."""
    B, D, H, W = masked.shape
    x = masked.permute(0, 2, 3, 1).reshape(B * H * W, D)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + "."))
    for i in range(n_layers):
        x = gate(x, sd[f"{prefix}.{i}.conv.weight"], sd[f"{prefix}.{i}.conv.bias"])
    return x.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()


def observed_points(mask_b: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(t, y, x) integer coordinates of mask>0 in lexicographic order.  mask_b: [T,H,W]."""
    tz, ty, tx = torch.nonzero(mask_b > 0, as_tuple=True)
    return tz, ty, tx


def idw_reference_style(points_xyz: Tensor, values: Tensor, shape: Tuple[int, int, int], k: int = 4,
                        tau: float = 0.05, chunk: int = 16384) -> Tensor:
    """IDW with the reference's own numerics: cdist (matmul path) + topk; tie order unspecified."""
    D, H, W = shape
    dev = points_xyz.device          # CPU in the tests; tools/gpu_eager_reference.py runs the same code with CUDA tensors
    gz, gy, gx = torch.meshgrid(torch.linspace(0, 1, D, device=dev), torch.linspace(0, 1, H, device=dev),
                                torch.linspace(0, 1, W, device=dev), indexing="ij")
    grid = torch.stack([gx, gy, gz], dim=-1).reshape(-1, 3).contiguous()
    out = torch.empty(grid.shape[0], dtype=torch.float32, device=dev)
    for s in range(0, grid.shape[0], chunk):
        e = min(s + chunk, grid.shape[0])
        d = torch.cdist(grid[s:e], points_xyz)
        dk, ik = torch.topk(d, k, dim=1, largest=False)
        inv = 1.0 / (dk + tau)
        w = inv * inv
        w = w / (w.sum(dim=1, keepdim=True) + 1e-12)
        out[s:e] = (values[ik] * w).sum(dim=1)
    return out.reshape(D, H, W)


def idw_integer_keys(tz: Tensor, ty: Tensor, tx: Tensor, shape: Tuple[int, int, int]) -> Tensor:
    """Exact squared distances scaled by ((W-1)(H-1)(D-1))^2 as int64: [Q, N].

    d^2 = (dx/(W-1))^2 + (dy/(H-1))^2 + (dt/(D-1))^2.  Multiplying by the common denominator
    gives an integer, so ordering and ties are exact.
    """
    D, H, W = shape
    sw, sh, sd = max(W - 1, 1), max(H - 1, 1), max(D - 1, 1)
    qz, qy, qx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    qz, qy, qx = qz.reshape(-1, 1), qy.reshape(-1, 1), qx.reshape(-1, 1)
    cx, cy, cz = (sh * sd) ** 2, (sw * sd) ** 2, (sw * sh) ** 2
    return cx * (qx - tx.reshape(1, -1)) ** 2 + cy * (qy - ty.reshape(1, -1)) ** 2 + cz * (qz - tz.reshape(1, -1)) ** 2


def idw_exact_table(tz: Tensor, ty: Tensor, tx: Tensor, shape: Tuple[int, int, int], k: int = 4, tau: float = 0.05,
                    chunk: int = 8192):
    """Neighbour indices [Q,kk] (int64 tensor) and normalised weights [Q,kk] (numpy float32) of ``idw_exact``: they depend on
    the point positions only, so samples that share a gauge pattern (the 'stis' mask) share the table."""
    D, H, W = shape
    Q, N = D * H * W, tz.numel()
    kk = min(k, N)
    sw, sh, sd = max(W - 1, 1), max(H - 1, 1), max(D - 1, 1)
    cx, cy, cz = (sh * sd) ** 2, (sw * sd) ** 2, (sw * sh) ** 2
    denom = float(sw * sh * sd)
    q = torch.arange(Q)
    qz, qy, qx = q // (H * W), (q // W) % H, q % W
    nb = torch.empty(Q, kk, dtype=torch.int64)
    wt = np.empty((Q, kk), dtype=np.float32)
    for s in range(0, Q, chunk):
        e = min(s + chunk, Q)
        key = (cx * (qx[s:e, None] - tx[None]) ** 2 + cy * (qy[s:e, None] - ty[None]) ** 2
               + cz * (qz[s:e, None] - tz[None]) ** 2)
        # stable ordering on (key, index): index < 2^20 always holds for our sizes
        comp = key * (1 << 20) + torch.arange(N)[None]
        ck, _ = torch.topk(comp, kk, dim=1, largest=False)
        nb[s:e] = ck & ((1 << 20) - 1)
        # The fp32 weight arithmetic runs in numpy (single-threaded): torch's multi-threaded CPU elementwise/reduction
        # path was observed to be irreproducible run-to-run (3e-5) on the GPU boxes' 16-core hosts.
        dk = np.sqrt((ck.numpy() >> 20).astype(np.float32)) / np.float32(denom)
        inv = np.float32(1.0) / (dk + np.float32(tau))
        w = inv * inv
        wt[s:e] = w / (w.sum(axis=1, keepdims=True, dtype=np.float32) + np.float32(1e-12))
    return nb, wt


def idw_exact(tz: Tensor, ty: Tensor, tx: Tensor, values: Tensor, shape: Tuple[int, int, int], k: int = 4,
              tau: float = 0.05, chunk: int = 8192, return_neighbors: bool = False, table=None):
    """IDW with exact integer ordering and the deterministic tie rule *smaller point index wins*.

    This is the behaviour the CUDA kernel implements (SURVEY.md 8c protocol).  Distances are
    evaluated in fp32 from the integer key: d = sqrt(key) / ((W-1)(H-1)(D-1)).
    ``table`` = a precomputed ``idw_exact_table`` of the same points (optional).
    """
    D, H, W = shape
    nb, wt = table if table is not None else idw_exact_table(tz, ty, tx, shape, k, tau, chunk)
    if values.requires_grad:       # autograd path (training-step oracle): linear in `values`
        out = (values[nb] * torch.from_numpy(wt)).sum(dim=1)
    else:
        values_np = values.detach().to(torch.float32).contiguous().numpy()
        out = torch.from_numpy((values_np[nb.numpy()] * wt).sum(axis=1, dtype=np.float32))
    out = out.reshape(D, H, W)
    return (out, nb) if return_neighbors else out


def idw_tie_mask(tz: Tensor, ty: Tensor, tx: Tensor, shape: Tuple[int, int, int], k: int = 4,
                 chunk: int = 8192) -> Tensor:
    """True where the k-th and (k+1)-th neighbour are exactly equidistant (ambiguous query)."""
    D, H, W = shape
    Q, N = D * H * W, tz.numel()
    if N <= k:
        return torch.zeros(D, H, W, dtype=torch.bool)
    sw, sh, sd = max(W - 1, 1), max(H - 1, 1), max(D - 1, 1)
    cx, cy, cz = (sh * sd) ** 2, (sw * sd) ** 2, (sw * sh) ** 2
    q = torch.arange(Q)
    qz, qy, qx = q // (H * W), (q // W) % H, q % W
    tie = torch.empty(Q, dtype=torch.bool)
    for s in range(0, Q, chunk):
        e = min(s + chunk, Q)
        key = (cx * (qx[s:e, None] - tx[None]) ** 2 + cy * (qy[s:e, None] - ty[None]) ** 2
               + cz * (qz[s:e, None] - tz[None]) ** 2)
        kk, _ = torch.topk(key, k + 1, dim=1, largest=False)
        tie[s:e] = kk[:, k - 1] == kk[:, k]
    return tie.reshape(D, H, W)


def idw_tie_candidates_ok(value: Tensor, tz, ty, tx, values: Tensor, shape, k: int = 4, tau: float = 0.05,
                          atol: float = 1e-4, max_queries: int = 4096) -> bool:
    """For tie queries: ``value`` must equal the IDW over *some* valid k-subset of the tied candidates."""
    import itertools
    D, H, W = shape
    tie = idw_tie_mask(tz, ty, tx, shape, k).reshape(-1)
    qs = torch.nonzero(tie).reshape(-1)[:max_queries]
    keys = idw_integer_keys(tz, ty, tx, shape)
    denom = float(max(W - 1, 1) * max(H - 1, 1) * max(D - 1, 1))
    flat = value.reshape(-1)
    for q in qs.tolist():
        row = keys[q]
        srt, order = torch.sort(row, stable=True)
        kth = srt[k - 1]
        fixed = order[srt < kth].tolist()
        tied = order[srt == kth].tolist()
        need = k - len(fixed)
        ok = False
        for combo in itertools.combinations(tied, need):
            sel = torch.tensor(fixed + list(combo))
            d = torch.sqrt(row[sel].double()) / denom
            w = 1.0 / (d + tau) ** 2
            w = w / (w.sum() + 1e-12)
            if abs(float((values[sel].double() * w).sum()) - float(flat[q])) <= atol:
                ok = True
                break
        if not ok:
            return False
    return True


def input_block(sd: State, masked: Tensor, masks: Tensor, k: int = 4, tau: float = 0.05,
                idw: str = "exact") -> Tensor:
    """InputBlock forward.  masked, masks: [B,16,H,W] -> [B,16,H,W] fp32."""
    B, D, H, W = masked.shape
    proc = gated_frames(sd, masked)
    outs = []
    tables = {}                      # neighbour tables of this call, keyed by the point pattern (shared by 'stis' batches)
    for b in range(B):
        tz, ty, tx = observed_points(masks[b])
        if tz.numel() == 0:
            outs.append(torch.zeros(D, H, W))
            continue
        vals = proc[b][tz, ty, tx]
        if idw == "exact":
            pk = (tz.numpy().tobytes(), ty.numpy().tobytes(), tx.numpy().tobytes())
            if pk not in tables:
                tables[pk] = idw_exact_table(tz, ty, tx, (D, H, W), k, tau)
            outs.append(idw_exact(tz, ty, tx, vals, (D, H, W), k, tau, table=tables[pk]))
        else:
            pts = torch.stack([tx.float() / max(W - 1, 1), ty.float() / max(H - 1, 1), tz.float() / max(D - 1, 1)],
                              dim=-1)
            outs.append(idw_reference_style(pts, vals, (D, H, W), k, tau))
    return torch.stack(outs, 0)


# --------------------------------------------------------------------------------------
# Generator trunk
# --------------------------------------------------------------------------------------


def doconv_weight(W: Tensor, D: Optional[Tensor], D_diag: Optional[Tensor], in_ch: int, out_ch: int,
                  groups: int, ksize: int) -> Tensor:
    """Effective conv weight of a DO-Conv layer: [out_ch, in_ch/groups, k, k]."""
    if ksize * ksize == 1:
        return W.reshape(out_ch, in_ch // groups, 1, 1)
    Wr = W.reshape(out_ch // groups, in_ch, W.shape[-1])          # raw reshape, as the reference does
    dow = torch.einsum("ims,ois->oim", D + D_diag, Wr)            # [out/g, in, k*k]
    return dow.reshape(out_ch, in_ch // groups, ksize, ksize)


def _do3x3(sd: State, key: str, x: Tensor, ch: int) -> Tensor:
    w = doconv_weight(sd[key + ".W"], sd[key + ".D"], sd[key + ".D_diag"], ch, ch, 1, 3)
    return F.conv2d(x, w, None, 1, 1)


def res_block(sd: State, prefix: str, x: Tensor) -> Tensor:
    ch = x.shape[1]
    y = torch.relu(_do3x3(sd, prefix + ".main.0.main.0", x, ch))
    y = _do3x3(sd, prefix + ".main.1.main.0", y, ch)
    return y + x


def e_block(sd: State, level: int, x: Tensor, num_res: int = 4) -> Tensor:
    for r in range(num_res):
        x = res_block(sd, f"Decoder.{level}.layers.{r}", x)
    return x


def pyramid_down(x: Tensor) -> Tensor:
    """max-pool 2x2 then every channel twice (out[c] = pooled[c // 2])."""
    p = F.max_pool2d(x, 2, 2)
    return p.repeat_interleave(2, dim=1)


def up_pos(sd: State, idx: int, x: Tensor) -> Tensor:
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    s = 2.0 * torch.sigmoid(sd[f"UP.{idx}.pos"])                  # 1 + (2*sigmoid - 1)
    x = x * s
    x = F.conv2d(x, sd[f"UP.{idx}.proj.weight"], sd[f"UP.{idx}.proj.bias"])
    return torch.relu(x)


def generator_trunk(sd: State, x: Tensor, return_intermediates: bool = False):
    """Everything after the InputBlock.  x: [B,16,H,W] -> pre-tanh z [B,16,H,W]."""
    inter: Dict[str, Tensor] = {}
    w_in = doconv_weight(sd["Convsin.0.main.0.W"], sd["Convsin.0.main.0.D"], sd["Convsin.0.main.0.D_diag"],
                         16, 64, 4, 3)
    x1 = F.conv2d(x, w_in, None, 1, 1, 1, 4) + x.repeat_interleave(4, dim=1)
    x2 = pyramid_down(x1)
    x4 = pyramid_down(x2)
    x8 = pyramid_down(x4)
    inter["stem"], inter["x4"], inter["x8"] = x1, x4, x8
    r = e_block(sd, 3, x8)
    inter["dec3"] = r
    r = up_pos(sd, 2, r)
    inter["up2"] = r
    r = e_block(sd, 2, x4 + r)
    inter["dec2"] = r
    r = up_pos(sd, 1, r)
    r = e_block(sd, 1, r)
    inter["dec1"] = r
    r = up_pos(sd, 0, r)
    r = e_block(sd, 0, r)
    inter["dec0"] = r
    w_out = doconv_weight(sd["ConvsOut.0.main.0.W"], None, None, 64, 16, 4, 1)
    z = F.conv2d(r, w_out, None, 1, 0, 1, 4)
    inter["z"] = z
    return (z, inter) if return_intermediates else z


def generator_forward(sd: State, masked_frames: Tensor, masks: Tensor, idw: str = "exact",
                      return_intermediates: bool = False):
    """P2IGenerator.forward.  [B,T,1,H,W] x2 -> [B,T,1,H,W]."""
    b, t, c, h, w = masked_frames.shape
    mf = masked_frames.reshape(b, c * t, h, w).float()
    mk = masks.reshape(b, c * t, h, w).float()
    x = input_block(sd, mf, mk, idw=idw)
    z, inter = generator_trunk(sd, x, True)
    inter["input"] = x
    out = torch.tanh(z).reshape(b, t, c, h, w)
    return (out, inter) if return_intermediates else out


# --------------------------------------------------------------------------------------
# Discriminator (dual branch, spectral norm)
# --------------------------------------------------------------------------------------

D2D_SPECS = [(0, 1), (2, 2), (4, 2), (6, 1), (8, 1)]                      # (seq index, stride)
D3D_SPECS = [(0, (1, 2, 2), 1), (2, (1, 2, 2), 1), (4, (1, 2, 2), 1), (6, (2, 1, 1), 1), (8, (1, 1, 1), 0)]


def _l2n(v: Tensor, eps: float = 1e-12) -> Tensor:
    return v / max(float(v.norm()), eps)


def spectral_weight(sd: Dict[str, Tensor], prefix: str, training: bool, update_state: bool = True) -> Tensor:
    """weight_orig / sigma with one power iteration when ``training`` (u, v updated in ``sd``)."""
    w = sd[prefix + ".weight_orig"]
    wm = w.reshape(w.shape[0], -1)
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    if training:
        with torch.no_grad():
            v = _l2n(wm.detach().t() @ u)
            u = _l2n(wm.detach() @ v)
        if update_state:
            sd[prefix + ".weight_u"], sd[prefix + ".weight_v"] = u.clone(), v.clone()
    sigma = torch.dot(u, wm @ v)
    return w / sigma


def discriminator_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = True, update_state: bool = True,
                          return_intermediates: bool = False):
    """P2IDiscriminator.forward.  x: [B,T,1,H,W] -> [B, (H/4)(W/4)]."""
    b, t, c, h, w = x.shape
    inter: Dict[str, Tensor] = {}
    y = x.reshape(b, t * c, h, w)
    for i, (idx, stride) in enumerate(D2D_SPECS):
        wgt = spectral_weight(sd, f"d2d.{idx}", training, update_state)
        y = F.conv2d(y, wgt, sd[f"d2d.{idx}.bias"], stride, 1)
        if i < len(D2D_SPECS) - 1:
            y = F.leaky_relu(y, 0.2)
        inter[f"d2d.{idx}"] = y
    z = x.permute(0, 2, 1, 3, 4)
    for i, (idx, stride, pad) in enumerate(D3D_SPECS):
        wgt = spectral_weight(sd, f"d3d.{idx}", training, update_state)
        z = F.conv3d(z, wgt, sd[f"d3d.{idx}.bias"], stride, pad)
        if i < len(D3D_SPECS) - 1:
            z = F.leaky_relu(z, 0.2)
        inter[f"d3d.{idx}"] = z
    z2 = z.mean(dim=2)
    if z2.shape[-2:] != y.shape[-2:]:
        z2 = F.interpolate(z2, size=y.shape[-2:], mode="bilinear", align_corners=False)
    fused = torch.sigmoid(sd["alpha2d"]) * y + z2
    out = fused.reshape(b, -1)
    return (out, inter) if return_intermediates else out


# --------------------------------------------------------------------------------------
# Losses
# --------------------------------------------------------------------------------------


def weighted_l1(pred: Tensor, target: Tensor) -> Tensor:
    wgt = 0.5 * torch.exp(5.14 * torch.clamp(target, max=0.7)) + 0.12
    # clamp(max=.7) equals the reference's where(y>0.7, w(0.7), w(y)) including at y == 0.7
    return (wgt * (pred - target).abs()).mean()


def temporal_kl(pred: Tensor, target: Tensor, temperature: float = 0.1) -> Tensor:
    """(1/B) * sum q (log q - log p); p, q = softmax over pixels of temporal differences."""
    dp = (pred[:, 1:] - pred[:, :-1]).flatten(2) / temperature
    dq = (target[:, 1:] - target[:, :-1]).flatten(2) / temperature
    p = torch.softmax(dp, dim=-1)
    q = torch.softmax(dq, dim=-1)
    return F.kl_div(p.log(), q, reduction="batchmean")


def reconstruction_loss(pred: Tensor, target: Tensor, k1_alpha: float = 0.0):
    pool = weighted_l1(pred, target)
    reg = temporal_kl(pred, target)
    return pool + k1_alpha * reg, {"pool": float(pool.detach()), "reg": float(reg.detach())}


def gan_loss(logits: Tensor, target_is_real: bool, loss_type: str = "nsgan", is_disc: Optional[bool] = False,
             target_real_label: float = 1.0, target_fake_label: float = 0.0) -> Tensor:
    if loss_type == "hinge":
        if is_disc is None:
            raise ValueError("`is_disc` must be set when using hinge loss.")
        if is_disc:
            return torch.relu(1 - logits).mean() if target_is_real else torch.relu(1 + logits).mean()
        return (-logits).mean()
    label = torch.full_like(logits, target_real_label if target_is_real else target_fake_label)
    if loss_type == "nsgan":
        return F.binary_cross_entropy(logits, label)
    if loss_type == "lsgan":
        return F.mse_loss(logits, label)
    raise ValueError(f"Unsupported GAN loss type: {loss_type}")


# --------------------------------------------------------------------------------------
# Metrics (streaming, sum-reduced states)
# --------------------------------------------------------------------------------------

EPS = 1e-10


def rain_rate(x: Tensor) -> Tensor:
    return torch.pow(10.0, x * 0.0625) * 0.036


def ssim_per_image(pred: Tensor, target: Tensor, data_range: float = 1.0, sigma: float = 1.5, k1: float = 0.01,
                   k2: float = 0.03) -> Tensor:
    """PARITY UNPINNED.  torchmetrics 1.0.3 (pinned by the reference's uv.lock:1162-1164; absent here) -- restatement of
    its published ``functional/image/ssim.py::_ssim_update`` as called by StructuralSimilarityIndexMeasure(data_range)
    (metric.py:36): gaussian window of size ``int(3.5*sigma+0.5)*2+1`` = 11, reflect padding by 5, grouped conv2d of
    (p, t, p*p, t*t, p*t), SSIM map, crop of the padded border, mean per image.  pred/target [N,1,H,W] f32 -> [N]."""
    ks = int(3.5 * sigma + 0.5) * 2 + 1
    pad = (ks - 1) // 2
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    dist = torch.arange((1 - ks) / 2, (1 + ks) / 2, 1, dtype=torch.float32)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    kernel = torch.matmul(g.t(), g).expand(1, 1, ks, ks)
    p = F.pad(pred, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    out = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel)
    n = pred.shape[0]
    mu_p, mu_t, e_pp, e_tt, e_pt = (out[i * n:(i + 1) * n] for i in range(5))
    mu_pp, mu_tt, mu_pt = mu_p * mu_p, mu_t * mu_t, mu_p * mu_t
    s_p, s_t, s_pt = e_pp - mu_pp, e_tt - mu_tt, e_pt - mu_pt
    full = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_p + s_t + c2))
    return full[..., pad:-pad, pad:-pad].reshape(n, -1).mean(-1)


class MetricSuiteOracle:
    """MAE/RMSE, POD/FAR/CSI/HSS per threshold, FSS per (threshold, scale), and SSIM ("ssim", parity unpinned:
    ``ssim_per_image``; only reported when ``with_ssim`` so that the reference goldens, taken with a stubbed
    torchmetrics, compare key for key)."""

    def __init__(self, thresholds: Sequence[float] = (0.5, 2.0, 4.0, 8.0), scales: Sequence[int] = (1, 2, 4, 8),
                 apply_transform: bool = True, with_ssim: bool = False, data_range: float = 1.0):
        self.with_ssim, self.data_range = with_ssim, data_range
        self.thr = [float(t) for t in thresholds]
        self.scales = [int(s) for s in scales]
        self.apply_transform = apply_transform
        self.reset()

    def reset(self):
        self.abs_sum = torch.tensor(0.0)
        self.sq_sum = torch.tensor(0.0)
        self.n = torch.tensor(0.0)
        self.cont = torch.zeros(len(self.thr), 4)                 # hits, misses, false, correct
        self.fss_sum = torch.zeros(len(self.thr), len(self.scales))
        self.fss_cnt = torch.zeros(len(self.thr), len(self.scales))
        self.ssim_sum, self.ssim_n = torch.tensor(0.0), 0

    def update(self, pred: Tensor, target: Tensor):
        p32, t32 = pred.detach().float(), target.detach().float()
        pr, tr = (rain_rate(p32), rain_rate(t32)) if self.apply_transform else (p32, t32)
        d = pr - tr
        self.abs_sum += d.abs().sum()
        self.sq_sum += (d * d).sum()
        self.n += d.numel()
        if self.with_ssim and pr.shape[-1] > 10 and pr.shape[-2] > 10:      # metric.py:55-56 (on the transformed values)
            sp = ssim_per_image(pr.reshape(-1, 1, *pr.shape[-2:]), tr.reshape(-1, 1, *tr.shape[-2:]), self.data_range)
            self.ssim_sum += sp.sum()
            self.ssim_n += sp.numel()
        pc, tc = rain_rate(p32), rain_rate(t32)                   # categorical/FSS always transform
        H, W = pc.shape[-2:]
        pm, tm = pc.reshape(-1, 1, H, W), tc.reshape(-1, 1, H, W)
        for i, thr in enumerate(self.thr):
            thr32 = torch.tensor(thr, dtype=torch.float32)
            a, o = pm >= thr32, tm >= thr32
            self.cont[i, 0] += (a & o).sum().float()
            self.cont[i, 1] += (~a & o).sum().float()
            self.cont[i, 2] += (a & ~o).sum().float()
            self.cont[i, 3] += (~a & ~o).sum().float()
            af, of = a.float(), o.float()
            for j, s in enumerate(self.scales):
                fa = F.avg_pool2d(af, s, 1, s // 2)
                fo = F.avg_pool2d(of, s, 1, s // 2)
                num = ((fa - fo) ** 2).mean()
                den = (fa * fa + fo * fo).mean()
                self.fss_sum[i, j] += 1.0 - num / (den + EPS)
                self.fss_cnt[i, j] += 1

    def compute(self) -> Dict[str, float]:
        n = torch.clamp(self.n, min=1.0)
        out = {"mae": float(self.abs_sum / n), "rmse": float(torch.sqrt(self.sq_sum / n))}
        if self.with_ssim:
            out["ssim"] = float(self.ssim_sum / self.ssim_n) if self.ssim_n else float("nan")
        for i, thr in enumerate(self.thr):
            h, m, f, c = self.cont[i]
            pre = f"cat_thr{thr:.2f}"
            out[f"{pre}/pod"] = float(h / (h + m + EPS))
            out[f"{pre}/far"] = float(f / (h + f + EPS))
            out[f"{pre}/csi"] = float(h / (h + m + f + EPS))
            den = (m + f) * (f + c) + (h + m) * (m + c)
            out[f"{pre}/hss"] = float(2 * (h * c - m * f) / (den + EPS))
        for i, thr in enumerate(self.thr):
            for j, s in enumerate(self.scales):
                if self.fss_cnt[i, j] > 0:
                    out[f"fss_thr{thr:.2f}_s{s}"] = float(self.fss_sum[i, j] / self.fss_cnt[i, j])
        return out


# --------------------------------------------------------------------------------------
# One full GAN training step (autograd on the functional oracle) and sliding-window inference
# --------------------------------------------------------------------------------------

G_FROZEN_SUFFIX = ".D_diag"


def adam_step(params: Dict[str, Tensor], grads: Dict[str, Tensor], state: Dict[str, Dict[str, Tensor]], step: int,
              lr: float = 1e-4, beta1: float = 0.0, beta2: float = 0.99, eps: float = 1e-8) -> None:
    """torch.optim.Adam (no weight decay, no amsgrad) on named tensors, in place."""
    for k, g in grads.items():
        if g is None:
            continue
        st = state.setdefault(k, {"m": torch.zeros_like(params[k]), "v": torch.zeros_like(params[k])})
        st["m"].mul_(beta1).add_(g, alpha=1 - beta1)
        st["v"].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(st["m"], denom, value=-lr / bc1)


def gan_train_step(g_sd: Dict[str, Tensor], d_sd: Dict[str, Tensor], frames: Tensor, masked: Tensor, masks: Tensor,
                   opt_g: Dict, opt_d: Dict, step: int, k1_weight: float = 0.05, adv_weight: float = 0.01,
                   lr: float = 1e-4, beta1: float = 0.0, beta2: float = 0.99, idw: str = "exact",
                   grads_out: Optional[Dict] = None) -> Dict[str, float]:
    """One iteration in the reference trainer's order (hinge loss).  Mutates g_sd/d_sd/optimizer state.
    grads_out (optional dict) receives {"g": {name: grad}, "d": {name: grad}, "preds": generator output} of this step."""
    g_train = [k for k in g_sd if not k.endswith(G_FROZEN_SUFFIX)]
    d_train = [k for k in d_sd if k.endswith("weight_orig") or k.endswith("bias") or k.startswith("alpha")]
    gp = {k: (g_sd[k].detach().clone().requires_grad_(True) if k in g_train else g_sd[k]) for k in g_sd}
    preds = generator_forward(gp, masked, masks, idw=idw)
    loss_g, parts = reconstruction_loss(preds, frames, k1_weight)

    dp = {k: (d_sd[k].detach().clone().requires_grad_(True) if k in d_train else d_sd[k]) for k in d_sd}
    lf = discriminator_forward(dp, preds.detach(), True)
    lr_ = discriminator_forward(dp, frames, True)
    loss_d = 0.5 * (gan_loss(lr_, True, "hinge", True) + gan_loss(lf, False, "hinge", True))
    used = [k for k in d_train if k != "alpha3d"]
    gd = torch.autograd.grad(loss_d, [dp[k] for k in used])
    for k in d_sd:                                                # carry the updated u/v back
        d_sd[k] = dp[k].detach() if k not in d_train else d_sd[k]
    adam_step(d_sd, dict(zip(used, gd)), opt_d, step, lr, beta1, beta2)

    dq = dict(d_sd)
    lg = discriminator_forward(dq, preds, True)
    for k in d_sd:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            d_sd[k] = dq[k].detach()
    adv = gan_loss(lg, True, "hinge", False) * adv_weight
    total = loss_g + adv
    gg = torch.autograd.grad(total, [gp[k] for k in g_train], allow_unused=True)
    if grads_out is not None:
        grads_out["g"], grads_out["d"], grads_out["preds"] = dict(zip(g_train, gg)), dict(zip(used, gd)), preds.detach()
    adam_step(g_sd, dict(zip(g_train, gg)), opt_g, step, lr, beta1, beta2)
    return {"rec": float(loss_g.detach()), "pool": parts["pool"], "reg": parts["reg"], "adv": float(adv.detach()),
            "dis": float(loss_d.detach()), "total": float(total.detach())}


def sliding_window_infer(sd: State, masked: Tensor, masks: Tensor, stride: int = 16, overlap: int = 12,
                         output_scale: float = 255.0, idw: str = "exact") -> Tensor:
    """Windowed generator inference of one event [1,L,1,H,W]: tail windows repeat the last frame,
    overlapping predictions are averaged, result scaled and clipped at 0.  Returns [L,1,H,W]."""
    L = masked.shape[1]
    step = max(1, stride - overlap)
    acc = torch.zeros(L, *masked.shape[2:])
    cnt = torch.zeros(L, 1, 1, 1)
    for s in range(0, L, step):
        e = s + stride
        mf, mk = masked[:, s:e], masks[:, s:e]
        valid = min(stride, L - s)
        if e > L:
            pad = e - L
            mf = torch.cat([mf, mf[:, -1:].expand(-1, pad, -1, -1, -1)], 1)
            mk = torch.cat([mk, mk[:, -1:].expand(-1, pad, -1, -1, -1)], 1)
        out = generator_forward(sd, mf, mk, idw=idw)
        acc[s:s + valid] += out[0, :valid]
        cnt[s:s + valid] += 1
    return torch.clamp(acc / torch.clamp(cnt, min=1e-5) * output_scale, min=0.0)


# ----------------------------------------------------------------------------------------------- batch preparation
def batch_post_process(video_u8, mask_u8, height: int, width: int):
    """numpy restatement of the host-side sample preparation (p2igan_bench/data/sti_dataset.py:203-243:
    ``astype(float32) / 255.0``, ``masked = video * mask``, ``_crop_center`` with start ``max((old - new) // 2, 0)``)
    followed by ``Trainer._prepare_batch``'s channel permute (scripts/train.py:468-473).
    video_u8 [B,T,H0,W0] uint8; mask_u8 [H0,W0] | [B,H0,W0] | [B,T,H0,W0] (non-zero = observed)
    -> (frames, masked, masks) float32 numpy arrays [B,T,1,height,width]."""
    import numpy as np
    v = np.asarray(video_u8)
    B, T, H0, W0 = v.shape
    m = np.asarray(mask_u8) != 0
    if m.ndim == 2:
        m = np.broadcast_to(m[None, None], v.shape)
    elif m.ndim == 3:
        m = np.broadcast_to(m[:, None], v.shape)
    frames = v.astype(np.float32) / np.float32(255.0)
    mask = m.astype(np.float32)
    masked = frames * mask
    y0, x0 = max((H0 - height) // 2, 0), max((W0 - width) // 2, 0)
    crop = lambda a: np.ascontiguousarray(a[:, :, y0:y0 + height, x0:x0 + width])[:, :, None]   # noqa: E731
    return crop(frames), crop(masked), crop(mask)
