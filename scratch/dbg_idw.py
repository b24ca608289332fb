import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch, synth
from oracle import p2i_oracle as O
from p2igan_b200 import build_generator, ops
DEV = "cuda:0"
H = W = 32
torch.manual_seed(3)
G = build_generator(synth.make_cfg(H, W))
with torch.no_grad():
    for l in G.input.layers:
        l.conv.bias.normal_(0, 0.1)
sd = {k: v.clone() for k, v in G.state_dict().items()}
B = 2
frames, masked, masks = synth.make_batch(B, 16, H, W, 12, 9, tie_free=False)
mf, mk = masked.reshape(B, 16, H, W), masks.reshape(B, 16, H, W)
ref = O.input_block(sd, mf, mk, idw="exact")
G = G.to(DEV)
outs = []
for it in range(int(os.environ.get("ITERS", "6"))):
    with torch.no_grad():
        out = G.input(mf.to(DEV), mk.to(DEV))
    torch.cuda.synchronize()
    d = (out.cpu() - ref).abs()
    outs.append(out.cpu())
    idx = torch.nonzero(d > 2e-5)
    print(it, float(d.max()), idx.shape[0], idx[:5].tolist(), "same as first:", torch.equal(outs[0], outs[-1]))
# examine one bad query
with torch.no_grad():
    pts, counts, src = ops.points_extract(mk.to(DEV).contiguous())
    print("counts", counts.tolist(), "src", src.tolist())
