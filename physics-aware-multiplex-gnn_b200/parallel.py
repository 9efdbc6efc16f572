"""Molecule-sharded data parallelism: one process per GPU, full parameter replica, each rank
runs its own batch; the only collective is ONE NCCL all-reduce over the flat gradient buffer per step.

The reference has no distributed code at all (single process, main_qm9.py:56-58); its loss is a mean over the
local batch (main_qm9.py:108), so gradients are averaged over ranks.
"""
import torch
import torch.distributed as dist


def flat_grad(model):
    """The module's flat gradient buffer if every live p.grad is a view into it (what our backward delivers
    after zero_grad), else None."""
    gflat, views = getattr(model, "_gflat", None), getattr(model, "_gviews", None)
    if gflat is None or views is None:
        return None
    for (_, p), v, used in zip(model._param_list, views, model._param_used):
        if used and p.requires_grad and p.grad is not v:
            return None
    return gflat


def allreduce_gradients(model, group=None, average=True):
    """All-reduce (mean) the gradients of `model` across the process group: one collective on the flat buffer
    when the grads alias it, otherwise a pack / all-reduce / unpack."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    buf = flat_grad(model)
    if buf is not None:
        if average and buf.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)      # averaged inside the collective: one launch
            return
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if average:
            buf.mul_(1.0 / world)
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    packed = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if average:
        packed.mul_(1.0 / world)
    off = 0
    for g in grads:
        g.copy_(packed[off:off + g.numel()].view_as(g))
        off += g.numel()


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (molecule sharding of a global batch)."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
