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


class OverlappedGradSync:
    """Bucketed gradient all-reduce that overlaps the backward pass.

    The flat gradient buffer is laid out per layer and backward finishes the layers last-to-first, so the slice of a
    layer is final long before the call's stream is.  The library records events at those points
    (pamnet_grad_buckets / pamnet_wait_grad_bucket); this class groups the 2L layer-halves into `n_buckets` contiguous
    slices and, right after ``loss.backward()`` has RETURNED (everything is enqueued, the GPU is still differentiating
    the early layers), issues one NCCL all-reduce per bucket on a communication stream that waits only for its bucket's
    events.  Only the last bucket (first layers + embeddings / basis MLPs, final when backward is) is exposed.

        sync = OverlappedGradSync(model)          # once; enables the bucket events
        loss.backward(); sync()                   # instead of allreduce_gradients(model)
        sync.wait()                               # before the optimizer step / reading p.grad

    ``allreduce_ms()`` returns the CUDA-event duration of the exposed tail of the last call (the time between the
    compute stream becoming idle and the last collective finishing) for the scaling report."""

    def __init__(self, model, group=None, n_buckets=4):
        from . import _lib
        self.model, self.group = model, group
        self.lib = _lib.load()
        _lib.check(self.lib.pamnet_grad_buckets(1), "grad_buckets")
        cfg = model._ccfg
        H = 2 * cfg.n_layer
        lo, hi = _lib.c_i64(), _lib.c_i64()
        spans = {}
        for h in range(H):
            _lib.check(self.lib.pamnet_grad_bucket_range(cfg, h, lo, hi), "grad_bucket_range")
            spans[h] = (lo.value, hi.value)
        # completion order: H-1 (last local layer) ... 0; the flat layout is [head | global 0..L-1 | local 0..L-1], so a run
        # of consecutive local (or global) layers, latest first, is one contiguous slice that is final with its LAST
        # member to complete (the lowest layer index of the run).
        L = cfg.n_layer
        per = max(1, (2 * L + max(1, n_buckets - 1) - 1) // max(1, n_buckets - 1) // 2)      # layers per bucket and kind
        self.buckets = []                           # (wait_half, lo, hi) in issue order
        for first in range(L - 1, -1, -per):        # layers first, first-1, ..., last
            last = max(0, first - per + 1)
            # local layers [last, first] complete with half 2*last+1; the global ones with half 2*last
            self.buckets.append((2 * last + 1, spans[2 * last + 1][0], spans[2 * first + 1][1]))
            self.buckets.append((2 * last, spans[2 * last][0], spans[2 * first][1]))
        # issue in completion order: a bucket whose gate half is larger completes earlier
        self.buckets.sort(key=lambda b: -b[0])
        self.head = (0, spans[0][0])                # embeddings, frequencies, basis MLPs: final with backward itself
        self.total = int(model._total)
        self.comm = None
        self._ev = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world == 1:
            return
        model = self.model
        buf = flat_grad(model)
        if buf is None:                             # grads do not alias the flat buffer: plain path
            allreduce_gradients(model, self.group)
            return
        dev = buf.device
        cur = torch.cuda.current_stream(dev)
        if self.comm is None or self.comm.device != dev:
            self.comm = torch.cuda.Stream(device=dev, priority=-1)
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        from . import _lib
        nccl = buf.is_cuda and dist.get_backend(self.group) == "nccl"
        op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        with torch.cuda.device(dev), torch.cuda.stream(self.comm):
            for half, lo, hi in self.buckets:
                _lib.check(self.lib.pamnet_wait_grad_bucket(half, self.comm.cuda_stream), "wait_grad_bucket")
                sl = buf[lo:hi]
                dist.all_reduce(sl, op=op, group=self.group)
                if not nccl:
                    sl.mul_(1.0 / self.world)
            # the head slice (and the compute stream's tail) last: this is the only exposed collective
            self._ev[0].record(cur)
            self.comm.wait_stream(cur)
            sl = buf[self.head[0]:self.head[1]]
            dist.all_reduce(sl, op=op, group=self.group)
            if not nccl:
                sl.mul_(1.0 / self.world)
            self._ev[1].record(self.comm)

    def wait(self):
        """Make the current stream wait for the collectives (call before the optimizer step)."""
        if self.world > 1 and self.comm is not None:
            torch.cuda.current_stream(self.comm.device).wait_stream(self.comm)

    def allreduce_ms(self):
        if self._ev is None:
            return 0.0
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])


class NativeGradSync:
    """Gradient averaging INSIDE ``loss.backward()``: the library owns an NCCL communicator (pamnet_comm_*) and
    model_backward enqueues one all-reduce per two-layer bucket on its communication stream as backward produces them
    (csrc/comm.cuh).  Nothing to call per step -- after ``backward()`` the gradients are the rank average.

        sync = NativeGradSync(model)      # once per process, after dist.init_process_group and torch.cuda.set_device
        ...
        sync.close()

    The unique id travels over the default torch.distributed group (any backend)."""

    def __init__(self, model, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.active = False
        if self.world == 1:
            return
        rank = dist.get_rank(group)
        dev = model._flat.device
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.check(self.lib.pamnet_comm_unique_id(buf), "comm_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        cbuf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
        with torch.cuda.device(dev):
            _lib.check(self.lib.pamnet_comm_init(cbuf, rank, self.world), "comm_init")
        self.dev = dev
        self.active = True

    def enable(self, on=True):
        if self.active:
            with torch.cuda.device(self.dev):
                self.lib.pamnet_comm_enable(1 if on else 0)

    def close(self):
        if self.active:
            with torch.cuda.device(self.dev):
                self.lib.pamnet_comm_destroy()
            self.active = False


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (molecule sharding of a global batch)."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
