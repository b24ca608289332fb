"""Diagnostic: run the InputBlock parity case in fresh processes and, on a mismatch, report which stage differs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)


def child():
    import torch
    import synth
    from oracle import p2i_oracle as O
    from p2igan_b200 import build_generator, ops
    H = W = 32
    torch.manual_seed(3)
    G = build_generator(synth.make_cfg(H, W))
    with torch.no_grad():
        for l in G.input.layers:
            l.conv.bias.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    B = 2
    frames, masked, masks = synth.make_batch(B, 16, H, W, 12, 9)
    mf, mk = masked.reshape(B, 16, H, W), masks.reshape(B, 16, H, W)
    ref = O.input_block(sd, mf, mk, idw="exact")
    proc = O.gated_frames(sd, mf)
    G = G.to("cuda:0")
    ib = G.input
    with torch.no_grad():
        out, ctx = ib.forward_ctx(mf.cuda(), mk.cuda(), save_for_backward=True)
    torch.cuda.synchronize()
    inp, pts, counts, src, table = ctx
    d = (out.cpu() - ref).abs()
    print(f"max {float(d.max()):.3e}")
    if float(d.max()) < 2e-5:
        return 0
    w0, b0, w1, b1 = (p.detach().contiguous() for p in ib.gate_params())
    vals, _ = ops.gate_points_fwd(inp, pts, counts, w0, b0, w1, b1)
    for b in range(B):
        tz, ty, tx = O.observed_points(mk[b])
        vref = proc[b][tz, ty, tx]
        n = int(counts[b])
        lin = (tz * H * W + ty * W + tx).int()
        print(f"b={b} n={n} pts equal {torch.equal(pts[b, :n].cpu(), lin)} vals maxdiff "
              f"{float((vals[b, :n].cpu() - vref).abs().max()):.3e} src={src.tolist()}")
        o_ref, nb = O.idw_exact(tz, ty, tx, vref, (16, H, W), return_neighbors=True)
        idx = table[0][int(src[b])].cpu().long()
        print("  neighbour sets differ at", int((idx.sort(1).values != nb.sort(1).values).any(1).sum()), "queries")
        bad = torch.nonzero((out[b].cpu() - o_ref).abs() >= 2e-5)
        print("  bad", bad.shape[0], bad[:6].tolist())
        if bad.shape[0]:
            q = int(bad[0][0]) * H * W + int(bad[0][1]) * W + int(bad[0][2])
            print("  q", q, "gpu idx", idx[q].tolist(), "w", table[1][int(src[b])][q].tolist(), "ref idx", nb[q].tolist(),
                  "gpu out", float(out[b].reshape(-1)[q]), "ref", float(o_ref.reshape(-1)[q]),
                  "vals gpu", vals[b][idx[q].cuda()].tolist(), "vals ref", vref[nb[q]].tolist())
    return 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sys.exit(child())
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    fails = 0
    for i in range(n):
        r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True)
        if r.returncode != 0:
            fails += 1
            print(f"--- run {i} rc={r.returncode}\n{r.stdout}\n{r.stderr[-2000:]}")
    print(f"{fails} failures of {n}")
