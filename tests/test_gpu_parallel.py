"""GPU, world_size 2 over NCCL (skipped with fewer than two devices): the data-parallel step equals the single-GPU step
on the concatenated batch -- for the one-collective path (allreduce_gradients), for
the bucketed all-reduce issued from Python (OverlappedGradSync) and for the library-owned NCCL communicator that averages
the buckets inside backward (NativeGradSync, csrc/comm.cu)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import pamnet_b200
        from pamnet_b200 import Config, PAMNet
        from pamnet_b200.data import synthetic_qm9_batch
        from pamnet_b200.parallel import NativeGradSync, OverlappedGradSync, allreduce_gradients
        torch.manual_seed(0)
        model = PAMNet(Config("QM9", 128, 3, 5.0, 5.0)).cuda()
        shards = [synthetic_qm9_batch(6, seed=10 + r) for r in range(world)]
        mine = shards[rank].to("cuda")

        def local_step():
            for p in model.parameters():
                p.grad = None
            out = model(mine)
            (out - mine.y).abs().mean().backward()        # per-rank MEAN, like main_qm9.py:108
            return out.detach()

        res = {}
        local_step()
        allreduce_gradients(model)                         # mean over ranks
        torch.cuda.synchronize()
        res["plain"] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
        sync = OverlappedGradSync(model, n_buckets=3)
        for _ in range(3):                                 # repeated: bucket events are re-recorded every backward
            local_step()
            sync()
            sync.wait()
        torch.cuda.synchronize()
        res["overlap"] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
        res["exposed_ms"] = sync.allreduce_ms()
        native = NativeGradSync(model)                     # library-owned communicator: averaged inside backward
        for _ in range(3):
            local_step()
        torch.cuda.synchronize()
        res["native"] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
        native.enable(False)
        local_step()
        torch.cuda.synchronize()
        res["native_off"] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
        native.close()
        if rank == 0:
            # single-GPU reference: the same shards, equal-size -> mean over ranks of per-rank means == mean over all graphs
            for p in model.parameters():
                p.grad = None
            for sh in shards:
                b = sh.to("cuda")
                out = model(b)
                ((out - b.y).abs().mean() / world).backward()        # accumulates
            torch.cuda.synchronize()
            res["single"] = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices (gpurun --gpus 2)")
def test_two_gpu_step_equals_single_gpu_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = res[0]["single"]
    scale = float(single.abs().max())
    for r in (0, 1):
        for kind in ("plain", "overlap", "native"):
            err = float((res[r][kind] - single).abs().max()) / scale
            assert err < 1e-5, (r, kind, err)                  # summation order of the collective / split-K atomics only
    assert torch.equal(res[0]["overlap"], res[1]["overlap"])    # both ranks hold the same averaged gradient
    assert torch.equal(res[0]["native"], res[1]["native"])
    assert not torch.equal(res[0]["native_off"], res[1]["native_off"])     # bypassed: each rank keeps its own gradient
