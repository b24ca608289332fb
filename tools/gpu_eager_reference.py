"""GPU-vs-GPU context number: what PyTorch eager + cuDNN/cuBLAS gives for the SAME
workloads on the same B200.  The reference checkout cannot travel to the GPU box, so this runs the oracle's restatement of
the reference (oracle/p2i_oracle.py: the same torch library calls -- F.conv2d/conv3d, cdist + topk IDW, einsum DO-Conv
composition, spectral-norm power iteration, autograd, Adam) with every tensor on cuda:0, fp32 storage, TF32 convolutions
allowed (PyTorch's default, as the reference runs them).  Not part of bench.py: the timed product path must not touch
cuDNN.  Writes one JSON object to stdout.

    python tools/gpu_eager_reference.py [--train-batch 16] [--infer-batch 32] [--steps 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "p2i-gan-benchmark_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import synth  # noqa: E402
from oracle import p2i_oracle as O  # noqa: E402
from p2igan_b200 import build_discriminator, build_generator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-batch", type=int, default=16)
    ap.add_argument("--infer-batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    cfg = synth.make_cfg(128, 128)
    torch.manual_seed(2024)
    g_sd = {k: v.detach().clone().to(dev) for k, v in build_generator(cfg).state_dict().items()}
    d_sd = {k: v.detach().clone().to(dev) for k, v in build_discriminator(cfg).state_dict().items()}
    out = {"what": "oracle restatement of the reference on cuda:0 (torch eager, cuDNN/cuBLAS, fp32 storage, TF32 convs allowed)",
           "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
           "allow_tf32_cudnn": torch.backends.cudnn.allow_tf32}

    def timed(fn, steps):
        fn(); fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    B = a.train_batch
    fr, mf, mk = (t.to(dev) for t in synth.make_batch(B, 16, 128, 128, 79, 1))
    og, od, it = {}, {}, [0]

    def train_step():
        it[0] += 1
        O.gan_train_step(g_sd, d_sd, fr, mf, mk, og, od, it[0], idw="ref")
    try:
        s = timed(train_step, a.steps)
        out["train"] = {"events_per_s": B / s, "ms_per_step": s * 1e3, "events_per_step": B}
    except Exception as exc:
        out["train"] = {"error": repr(exc)}
    B = a.infer_batch
    fr, mf, mk = (t.to(dev) for t in synth.make_batch(B, 16, 128, 128, 79, 1))

    def infer_step():
        with torch.no_grad():
            O.generator_forward(g_sd, mf, mk, idw="ref")
    try:
        s = timed(infer_step, a.steps)
        out["infer"] = {"events_per_s": B / s, "ms_per_step": s * 1e3, "events_per_step": B}
    except Exception as exc:
        out["infer"] = {"error": repr(exc)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
