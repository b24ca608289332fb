"""CPU emulation of the lane-level index maps of csrc/d3d_first_mma.cu (development aid, no GPU needed): every shared-memory /
global address and every mma.sync fragment slot is computed with the kernel's formulas in numpy, the MMA itself from the PTX
fragment definition, and the result is compared with torch's conv3d / its autograd.  Values stay fp32 here (the kernel rounds
operands to bf16), so agreement is to fp32 round-off: what is checked is the INDEXING."""
import numpy as np
import torch
import torch.nn.functional as F


def tap_offset(k, WP):
    if k >= 27:
        return -1
    kt, r = divmod(k, 9)
    ky, kx = divmod(r, 3)
    return (kt * 3 + ky) * WP + kx


def stage_rows(x, b, yo, T, H, W, WP):
    xs = np.full(((T + 2) * 3 * WP,), np.nan, np.float32)
    W4 = W // 4
    for i in range((T + 2) * 3 * W4):
        c4, fr = i % W4, i // W4
        f, r = divmod(fr, 3)
        ti, yi = f - 1, 2 * yo + r - 1
        v = np.zeros(4, np.float32)
        if 0 <= ti < T and 0 <= yi < H:
            v = x[b, ti, yi, 4 * c4:4 * c4 + 4]
        xs[fr * WP + 4 * c4 + 1: fr * WP + 4 * c4 + 5] = v
    for i in range((T + 2) * 3):
        xs[i * WP] = 0
        xs[i * WP + W + 1] = 0
    return xs


def mma(acc, a, b0, b1):
    """acc[lane][4], a[lane][4][2], b0/b1[lane][2] -> acc (m16n8k16 fragment maps)."""
    A = np.zeros((16, 16), np.float64); Bm = np.zeros((16, 8), np.float64); C = np.zeros((16, 8), np.float64)
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * q + e] = a[lane][0][e]; A[g + 8, 2 * q + e] = a[lane][1][e]
            A[g, 2 * q + 8 + e] = a[lane][2][e]; A[g + 8, 2 * q + 8 + e] = a[lane][3][e]
            Bm[2 * q + e, g] = b0[lane][e]; Bm[2 * q + 8 + e, g] = b1[lane][e]
            C[g, 2 * q + e] = acc[lane][e]; C[g + 8, 2 * q + e] = acc[lane][2 + e]
    D = A @ Bm + C
    out = np.zeros((32, 4))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        out[lane] = [D[g, 2 * q], D[g, 2 * q + 1], D[g + 8, 2 * q], D[g + 8, 2 * q + 1]]
    return out


def fwd(x, w, bias, B, T, H, W):
    WP, Ho, Wo = W + 2, H // 2, W // 2
    y = np.full((B * T * (Ho // 2) * (Wo // 2) * 4 * 32,), np.nan, np.float32)
    for blk in range(B * Ho):
        b, yo = divmod(blk, Ho)
        xs = stage_rows(x, b, yo, T, H, W, WP)
        tiles_x = Wo >> 4
        for tile in range(T * tiles_x):
            t, xt = divmod(tile, tiles_x)
            x0 = xt << 4
            stage = np.full((16 * 40,), np.nan, np.float32)           # 80-byte pitch = 40 bf16 slots
            acc = [[None] * 4 for _ in range(4)]
            a = [[[[0, 0] for _ in range(4)] for _ in range(32)] for _ in range(2)]
            bw = [[[[[0, 0] for _ in range(2)] for _ in range(32)] for _ in range(4)] for _ in range(2)]
            for lane in range(32):
                g, q = lane >> 2, lane & 3
                base = t * 3 * WP + 2 * (x0 + g)
                for s in range(2):
                    for h in range(2):
                        k0 = 16 * s + 8 * h + 2 * q
                        o = [tap_offset(k0, WP), tap_offset(k0 + 1, WP)]
                        for e in range(2):
                            a[s][lane][2 * h][e] = xs[base + o[e]] if o[e] >= 0 else 0.0
                            a[s][lane][2 * h + 1][e] = xs[base + o[e] + 16] if o[e] >= 0 else 0.0
                            for j in range(4):
                                n = 8 * j + g
                                bw[s][j][lane][h][e] = w[n, k0 + e] if k0 + e < 27 else 0.0
            for j in range(4):
                accj = np.zeros((32, 4))
                for lane in range(32):
                    q = lane & 3
                    accj[lane] = [bias[8 * j + 2 * q], bias[8 * j + 2 * q + 1]] * 2
                for s in range(2):
                    accj = mma(accj, a[s], [bw[s][j][l][0] for l in range(32)], [bw[s][j][l][1] for l in range(32)])
                for lane in range(32):
                    g, q = lane >> 2, lane & 3
                    v = np.where(accj[lane] > 0, accj[lane], 0.2 * accj[lane])
                    stage[g * 40 + 8 * j + 2 * q: g * 40 + 8 * j + 2 * q + 2] = v[:2]
                    stage[(g + 8) * 40 + 8 * j + 2 * q: (g + 8) * 40 + 8 * j + 2 * q + 2] = v[2:]
            for it in range(2):
                for lane in range(32):
                    pix, chunk = (lane >> 2) + 8 * it, lane & 3
                    v = stage[pix * 40 + chunk * 8: pix * 40 + chunk * 8 + 8]
                    xo = x0 + pix
                    o = ((((b * T + t) * (Ho >> 1) + (yo >> 1)) * (Wo >> 1) + (xo >> 1)) * 4 + ((yo & 1) * 2 + (xo & 1))) * 32 + chunk * 8
                    y[o:o + 8] = v
    return y


def bwd_w(dpre, x, B, T, H, W):
    WP, Ho, Wo = W + 2, H // 2, W // 2
    dW = np.zeros((32, 27)); db = np.zeros(32)
    dflat = dpre.reshape(-1)
    for blk in range(B * Ho):
        b, yo = divmod(blk, Ho)
        xs = stage_rows(x, b, yo, T, H, W, WP)
        red = np.zeros((32, 33))
        tiles_x = Wo >> 4
        acc = [[np.zeros((32, 4)) for _ in range(4)] for _ in range(2)]
        for tile in range(T * tiles_x):
            t, xt = divmod(tile, tiles_x)
            x0 = xt << 4
            src = (((b * T + t) * Ho + yo) * Wo + x0) * 32
            stage = np.full((16 * 40,), np.nan, np.float32)
            for it in range(2):
                for lane in range(32):
                    idx = lane + 32 * it
                    stage[(idx >> 2) * 40 + (idx & 3) * 8: (idx >> 2) * 40 + (idx & 3) * 8 + 8] = dflat[src + idx * 8: src + idx * 8 + 8]
            # ldmatrix.x4.trans: lane l supplies the row address of matrix l>>3, row l&7; thread (g,q) gets M_i[2q][g], M_i[2q+1][g]
            a = [[[[0, 0] for _ in range(4)] for _ in range(32)] for _ in range(2)]
            for m in range(2):
                mats = []
                for i in range(4):
                    rows = []
                    for r in range(8):
                        l = 8 * i + r
                        addr = (8 * (l >> 4) + (l & 7)) * 40 + ((l >> 3) & 1) * 8 + 16 * m      # in bf16 slots
                        rows.append(stage[addr:addr + 8])
                    mats.append(np.stack(rows))
                for lane in range(32):
                    g, q = lane >> 2, lane & 3
                    for i in range(4):
                        a[m][lane][i] = [mats[i][2 * q][g], mats[i][2 * q + 1][g]]
            base = t * 3 * WP + 2 * x0
            for j in range(4):
                b0 = [[0, 0] for _ in range(32)]; b1 = [[0, 0] for _ in range(32)]
                for lane in range(32):
                    g, q = lane >> 2, lane & 3
                    o = tap_offset(8 * j + g, WP)
                    if o >= 0:
                        p = base + o + 4 * q
                        b0[lane] = [xs[p], xs[p + 2]]; b1[lane] = [xs[p + 16], xs[p + 18]]
                    elif j == 3 and g == 3:
                        b0[lane] = [1.0, 1.0]; b1[lane] = [1.0, 1.0]
                for m in range(2):
                    acc[m][j] = mma(acc[m][j], a[m], b0, b1)
        for m in range(2):
            for j in range(4):
                for lane in range(32):
                    g, q = lane >> 2, lane & 3
                    c0, k0 = 16 * m + g, 8 * j + 2 * q
                    red[c0, k0] += acc[m][j][lane][0]; red[c0, k0 + 1] += acc[m][j][lane][1]
                    red[c0 + 8, k0] += acc[m][j][lane][2]; red[c0 + 8, k0 + 1] += acc[m][j][lane][3]
        dW += red[:, :27]; db += red[:, 27]
    return dW, db


def run(B=2, T=3, H=8, W=32, seed=0):
    """-> (max |forward - torch|, max |dW - autograd|, max |db - autograd|, number of unwritten output slots)."""
    torch.manual_seed(seed)
    x = torch.randn(B, T, H, W)
    w = torch.randn(32, 1, 3, 3, 3, requires_grad=True)
    bias = torch.randn(32, requires_grad=True)
    pre = F.conv3d(x[:, None], w, bias, stride=(1, 2, 2), padding=1)
    ref = F.leaky_relu(pre, 0.2).detach()                                    # [B, 32, T, Ho, Wo]
    y = fwd(x.numpy(), w.detach().reshape(32, 27).numpy(), bias.detach().numpy(), B, T, H, W)
    Ho, Wo = H // 2, W // 2
    y = torch.from_numpy(y).reshape(B, T, Ho // 2, Wo // 2, 2, 2, 32).permute(0, 6, 1, 2, 4, 3, 5).reshape(B, 32, T, Ho, Wo)
    dpre = torch.randn_like(pre)
    pre.backward(dpre)
    dW, db = bwd_w(dpre.permute(0, 2, 3, 4, 1).contiguous().numpy(), x.numpy(), B, T, H, W)
    return (float((y - ref).abs().max()), float(np.abs(dW - w.grad.reshape(32, 27).numpy()).max()),
            float(np.abs(db - bias.grad.numpy()).max()), int(torch.isnan(y).sum()))


if __name__ == "__main__":
    e_fwd, e_dw, e_db, holes = run()
    print("fwd max abs diff", e_fwd, "unwritten", holes)
    print("dW max abs diff", e_dw, "db", e_db)
