"""CPU, world_size 2, gloo: the data-parallel host logic of the training step --
flat gradient buffer, one all-reduce(sum) per model, 1/world folded into the optimiser -- and the
equivalence  N ranks x b events == 1 rank x N*b events  on a small model."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "p2i-gan-benchmark_b200"))
    from p2igan_b200.train_step import FlatGrads
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
        flat = FlatGrads(list(model.parameters()) + [frozen])
        assert frozen.grad is None and len(flat.params) == 4
        assert flat.flat.numel() == sum(p.numel() for p in model.parameters())
        for p in model.parameters():                        # p.grad are views of the flat buffer
            assert p.grad.untyped_storage().data_ptr() == flat.flat.untyped_storage().data_ptr()
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 6, generator=g)
        y = torch.randn(8, 1, generator=g)
        xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]   # this rank's shard
        flat.zero()
        ((model(xs) - ys) ** 2).mean().backward()           # accumulates into the flat views
        scale = flat.all_reduce()
        assert scale == 1.0 / world
        dp_grad = flat.flat.clone() * scale
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        ref.load_state_dict(model.state_dict())
        ((ref(x) - y) ** 2).mean().backward()               # single process, global batch
        ref_grad = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        q.put((rank, float((dp_grad - ref_grad).abs().max()), None))
    except Exception as e:  # noqa: BLE001
        q.put((rank, None, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_grads_allreduce_equals_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, exc in res:
        assert exc is None, exc
        assert err < 1e-6, (rank, err)


def test_flat_grads_single_process_noop():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "p2i-gan-benchmark_b200"))
    from p2igan_b200.train_step import FlatGrads
    p = torch.nn.Parameter(torch.zeros(4))
    f = FlatGrads([p])
    p.grad.add_(1.0)
    assert f.all_reduce() == 1.0 and float(f.flat.sum()) == 4.0
    f.zero()
    assert float(p.grad.abs().sum()) == 0.0
