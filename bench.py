#!/usr/bin/env python
"""bench.py -- throughput of the P2I-GAN hot path on B200 (contract: see the task's bench section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train] [--impl ours|reference]

N=1 default workload: BASELINE.json configs[1] -- generator-only inference, batch 32 synthetic
radar-input events of 16x128x128, random-init weights (seed 2024), 79 gauge pixels.
One "step" = one pass of the hot path over one batch.  Prints ONE JSON line on rank 0.

  value        events/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          events/s through the public module API with pinned HOST buffers: H2D of the step's
               inputs and D2H of the step's result inside the timed region
  roofline     dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs / CUDA-event time
               of its launches, against MEASURED_PEAKS.json
  cpu_baseline the oracle (CPU restatement of the reference, oracle/) timed on this box's host cores
               on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "p2i-gan-benchmark_b200"))

import torch  # noqa: E402

import synth  # noqa: E402

T, H, W = 16, 128, 128
N_OBS = 79
G_FWD_FLOP_PER_EVENT = 39.54e9       # SURVEY.md 8d (2*MAC of the reference's conv calls)


def conv_flops_fwd(B):
    """Algorithmic FLOPs (2*MAC) of the tensor-core kernel launches in one generator forward, as executed
    (the UPPos projection runs at low resolution: 4x fewer MACs than the reference's)."""
    f = 0.0
    for lvl, C in enumerate((64, 128, 256, 512)):
        hw = (H >> lvl) * (W >> lvl)
        f += 8 * 2.0 * B * hw * C * C * 9
    for lvl, C in ((1, 128), (2, 256), (3, 512)):
        hw = (H >> lvl) * (W >> lvl)
        f += 2.0 * B * hw * C * (C // 2)
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[1]))
                mx = float(p[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1382.0), d.get("bf16_tflops", 1607.2), d.get("hbm_gbs", 6547.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------- CPU / reference arm
def cpu_generator_events_per_s(n_events: int, steps: int, warmup: int):
    """Oracle (CPU restatement of the reference generator, reference-style IDW numerics) on the host cores."""
    from oracle import p2i_oracle as O
    from p2igan_b200 import build_generator
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    torch.manual_seed(2024)
    sd = {k: v.detach().clone() for k, v in build_generator(synth.make_cfg(H, W)).state_dict().items()}
    frames, masked, masks = synth.make_batch(n_events, T, H, W, N_OBS, 1)
    with torch.no_grad():
        for _ in range(warmup):
            O.generator_forward(sd, masked, masks, idw="ref")
        t0 = time.perf_counter()
        for _ in range(steps):
            O.generator_forward(sd, masked, masks, idw="ref")
        dt = time.perf_counter() - t0
    return n_events * steps / dt, cores, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_events = 2
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    v, cores, spp = cpu_generator_events_per_s(n_events, steps, warmup)
    line = {"impl": "reference", "metric": "infer events/s (generator forward, 16x128x128)", "value": v, "unit": "events/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spp * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "P2IGAN generator-only inference, synthetic radar-input events 16x128x128 (BASELINE configs[1])",
                       "events_per_step": n_events, "gauge_pixels": N_OBS},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": cores, "kind": "port",
                             "sample": f"{n_events} events/step x {steps} steps of the same workload (oracle/p2i_oracle.py, "
                                       "torch CPU fp32, reference-style cdist/topk IDW)"},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from p2igan_b200 import build_generator, ops
    from p2igan_b200._lib import LIB

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch

    torch.manual_seed(2024)
    G = build_generator(synth.make_cfg(H, W)).to(dev).eval()
    # four rotating input batches (per-rank seeds) so no step re-reads the previous step's inputs from L2
    batches = []
    for i in range(4):
        frames, masked, masks = synth.make_batch(B, T, H, W, N_OBS, 1000 * rank + i)
        batches.append((masked.to(dev), masks.to(dev)))
    host = [(m.cpu().pin_memory(), k.cpu().pin_memory()) for m, k in batches[:2]]
    out_host = torch.empty(B, T, 1, H, W, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        with torch.no_grad():
            return G(*batches[i % 4])

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = LIB.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = LIB.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- e2e: public API, pinned host inputs -> H2D, forward, D2H of the prediction, every step
    def e2e_step(i):
        m, k = host[i % 2]
        with torch.no_grad():
            o = G(m.to(dev, non_blocking=True), k.to(dev, non_blocking=True))
        out_host.copy_(o, non_blocking=True)

    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- dominant kernel: events around every tcgen05 conv launch of the same steps
    conv_ms = 0.0
    n_conv = 0
    orig = ops.conv2d_cl
    evs = []

    def timed_conv(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(*a, **k)
        e.record()
        evs.append((s, e))
        return r

    ops.conv2d_cl = timed_conv
    import p2igan_b200.layers as _layers
    import p2igan_b200.generator as _gen
    psteps = min(args.steps, 5)
    for i in range(psteps):
        step(i)
    torch.cuda.synchronize()
    ops.conv2d_cl = orig
    for s, e in evs:
        conv_ms += s.elapsed_time(e)
    n_conv = len(evs)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        sustained, burst, hbm, src = peaks()
        events = B * world * args.steps
        value = events / (ms * 1e-3)
        conv_tflops = (conv_flops_fwd(B) * psteps / (conv_ms * 1e-3)) / 1e12 if conv_ms > 0 else None
        line = {
            "metric": "infer events/s (generator forward, 16x128x128)", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "P2IGAN generator-only inference, batch 32 synthetic radar-input events 16x128x128 "
                                   "(BASELINE configs[1])", "events_per_step_per_gpu": B, "gauge_pixels": N_OBS,
                       "weights": "random init seed 2024", "parallelism": f"events sharded over {world} GPU(s), no collective",
                       "l2": "no explicit flush: 4 rotating input batches and a per-step working set (~2.6 GB of "
                             "activations at B=32) far above the 126 MB L2"},
            "e2e": {"value": events / (ms_e2e * 1e-3), "unit": "events/s",
                    "h2d_bytes_per_step": 2 * B * T * H * W * 4, "d2h_bytes_per_step": B * T * H * W * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 implicit GEMM, 35 launches/step)",
                         "achieved": conv_tflops, "peak": sustained, "unit": "TFLOP/s",
                         "frac": (conv_tflops / sustained) if conv_tflops else None, "traffic": None,
                         "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                         "kernel_ms_per_step": conv_ms / max(psteps, 1), "launches_timed": n_conv},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            v, cores, spp = cpu_generator_events_per_s(2, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "events/s", "cores": cores, "kind": "port",
                                    "sample": "2 events/step x 3 steps of the same workload (oracle/p2i_oracle.py, torch CPU "
                                              "fp32, reference-style IDW)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="events per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
