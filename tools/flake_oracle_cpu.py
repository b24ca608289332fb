"""Diagnostic (CPU only): is the oracle's InputBlock reproducible inside one fresh process on this host?"""
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
    from p2igan_b200 import build_generator
    H = W = 32
    torch.manual_seed(3)
    G = build_generator(synth.make_cfg(H, W))
    with torch.no_grad():
        for l in G.input.layers:
            l.conv.bias.normal_(0, 0.1)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    frames, masked, masks = synth.make_batch(2, 16, H, W, 12, 9)
    mf, mk = masked.reshape(2, 16, H, W), masks.reshape(2, 16, H, W)
    p = [O.gated_frames(sd, mf) for _ in range(3)]
    tz, ty, tx = O.observed_points(mk[0])
    v = p[2][0][tz, ty, tx]
    r = [O.idw_exact(tz, ty, tx, v, (16, H, W), return_neighbors=True) for _ in range(3)]
    dp = max(float((p[i] - p[2]).abs().max()) for i in range(2))
    dr = max(float((r[i][0] - r[2][0]).abs().max()) for i in range(2))
    dn = max(int((r[i][1] != r[2][1]).sum()) for i in range(2))
    print(f"threads {torch.get_num_threads()} gate diff {dp:.3e} idw diff {dr:.3e} nbr diff {dn}")
    return 1 if (dp > 0 or dr > 0 or dn > 0) else 0


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        sys.exit(child())
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    fails = 0
    for i in range(n):
        r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True)
        if r.returncode != 0:
            fails += 1
            print(f"run {i}: {r.stdout.strip()} {r.stderr[-300:]}")
    print(f"{fails} irreproducible of {n}")
