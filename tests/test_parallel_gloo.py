"""CPU, world_size 2 over gloo: the one collective of the data-parallel path (flat-gradient all-reduce) and
the molecule sharding helper."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import pamnet_b200
        from pamnet_b200 import Config, PAMNet
        from pamnet_b200.parallel import allreduce_gradients, flat_grad
        torch.manual_seed(0)
        model = PAMNet(Config("QM9", 16, 1, 5.0, 5.0))
        # (a) grads that alias one flat buffer, the layout our backward produces
        views = model._grad_views()
        model._gflat.fill_(float(rank + 1))
        for (name, p), v, used in zip(model._param_list, views, model._param_used):
            p.grad = v if used else None
        assert flat_grad(model) is not None
        allreduce_gradients(model)
        ok_a = all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in model.parameters() if p.grad is not None)
        # (b) independent grad tensors -> pack / reduce / unpack
        for p in model.parameters():
            p.grad = torch.full_like(p, float(10 * (rank + 1)))
        assert flat_grad(model) is None
        allreduce_gradients(model)
        ok_b = all(torch.allclose(p.grad, torch.full_like(p.grad, 15.0)) for p in model.parameters())
        q.put((rank, ok_a, ok_b))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)


def test_shard_range_partitions_molecules():
    from pamnet_b200.parallel import shard_range
    for n, w in [(32, 8), (33, 8), (5, 8), (256, 3)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
