import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch, synth
from oracle import p2i_oracle as O
from p2igan_b200 import build_generator
H=W=32; B=2
torch.manual_seed(2024)
G = build_generator(synth.make_cfg(H, W))
gen = torch.Generator().manual_seed(11)
with torch.no_grad():
    for n, p in G.named_parameters():
        if n.endswith(".D") or n.endswith(".pos") or n.endswith("proj.bias") or n.endswith("conv.bias"):
            p.add_(torch.randn(p.shape, generator=gen) * 0.05)
sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
frames, masked, masks = synth.make_batch(B, 16, H, W, 12, 3)
def grads(autocast):
    train = [k for k in sd if not k.endswith(".D_diag")]
    p = {k: (v.clone().requires_grad_(True) if k in train else v) for k, v in sd.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = O.generator_forward(p, masked, masks, idw="exact")
    loss, _ = O.reconstruction_loss(out.float(), frames, 0.05)
    g = torch.autograd.grad(loss, [p[k] for k in train], allow_unused=True)
    return dict(zip(train, g))
g32 = grads(False); g16 = grads(True)
rs = []
for k in g32:
    if g32[k] is None: continue
    r = float((g16[k].double()-g32[k].double()).norm()/g32[k].double().norm().clamp_min(1e-20))
    rs.append((r,k))
rs.sort(reverse=True)
print(rs[:10]); print("median", rs[len(rs)//2])
