"""``Config`` / ``PAMNet`` / ``PAMNet_s`` with the reference's constructor signatures, module tree and
state_dict keys (models.py:12-56, :227-258), executing on libpamnet_sm100.so.

Drop-in for ``from models import PAMNet, PAMNet_s, Config`` (main_qm9.py:13): ``model(data)`` returns
Tensor[num_graphs] on the data's device, differentiable w.r.t. every parameter (one autograd node whose
backward is the hand-written CUDA backward).  CUDA only; there is no CPU / eager fallback.
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from .basis import sbf_consts_struct
from .layers import (MLP, BesselBasisLayer, Global_MessagePassing, Local_MessagePassing,
                     Local_MessagePassing_s, SphericalBasisLayer)

_DATASET_IDS = {"QM9": 0, "PDBbind": 1}


class Config(object):
    def __init__(self, dataset, dim, n_layer, cutoff_l, cutoff_g, flow='source_to_target'):
        self.dataset = dataset
        self.dim = dim
        self.n_layer = n_layer
        self.cutoff_l = cutoff_l
        self.cutoff_g = cutoff_g
        self.flow = flow


def _dataset_kind(name):
    if name[:3].lower() == "rna":
        return 2
    return _DATASET_IDS.get(name, -1)


class GraphPlan:
    """Per-batch execution plan (device blobs + sizes); layer-invariant (models.py:104-188)."""

    def __init__(self, sizes, base, trip, edge_index_g, edge_index_l):
        self.sizes, self.base, self.trip = sizes, base, trip
        self.edge_index_g, self.edge_index_l = edge_index_g, edge_index_l

    _ARRAYS = ("g_ptr", "g_src", "g_eid", "l_ptr", "l_src", "l_dst", "l_eid", "t_ptr", "t_gather", "t_owner", "t_split",
               "dist_g", "dist_l", "t_angle", "g_dst", "g_optr", "g_opos", "l_optr", "l_opos", "t_cnt", "tt_ptr", "tt_t",
               "n2g", "gptr")

    def arrays(self):
        """The plan's arrays as a dict of tensors (views of the two blobs) -- test / debugging aid; names and meaning:
        csrc/graph.cuh:Plan."""
        lib, sz = _lib.load(), self.sizes
        n, g, eg, el, t = sz.n_nodes, sz.n_graphs, sz.n_edges_g, sz.n_edges_l, sz.n_t2 + sz.n_t1
        count = {"g_ptr": n + 1, "g_optr": n + 1, "l_ptr": n + 1, "l_optr": n + 1, "n2g": n, "gptr": g + 1,
                 "g_src": eg, "g_dst": eg, "g_eid": eg, "g_opos": eg, "dist_g": eg,
                 "l_src": el, "l_dst": el, "l_eid": el, "l_opos": el, "dist_l": el, "t_split": el, "t_cnt": el,
                 "t_ptr": el + 1, "tt_ptr": el + 1, "t_gather": t, "t_owner": t, "tt_t": t, "t_angle": t}
        out = {}
        for which, name in enumerate(self._ARRAYS):
            in_trip = _lib.c_i32()
            off = lib.pamnet_debug_plan_offset(sz, which, in_trip)
            if off < 0:
                raise _lib.PamnetError("plan array %s not available" % name)
            blob = self.trip if in_trip.value else self.base
            dtype = torch.float32 if name in ("dist_g", "dist_l", "t_angle") else torch.int32
            out[name] = blob[off:off + 4 * count[name]].view(dtype)
        return out


class _PAMNetFunction(torch.autograd.Function):
    """One autograd node for the whole model.  Its only differentiable input is the flat parameter buffer; the
    backward writes the hand-written CUDA gradients straight into the module's flat gradient buffer and
    (re)attaches ``p.grad`` views with the usual accumulate semantics (see _PAMNetBase._deliver_grads) --
    routing 390 tensors through autograd costs more host time than the whole GPU step."""

    @staticmethod
    def forward(ctx, mod, plan, node_in, sign, pos, flat, prepared):
        lib = _lib.load()
        dev = pos.device
        cfg, sz = mod._ccfg, plan.sizes
        ws_bytes = lib.pamnet_workspace_bytes(cfg, sz)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty(sz.n_graphs, dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[5]
        # the library launches on the CURRENT device and owns per-device streams: make the tensors' device current
        # (a model on cuda:1 while cuda:0 is current would otherwise run on device 0 with device-1 pointers)
        with torch.cuda.device(dev):
            _lib.check(lib.pamnet_model_forward(cfg, sz, sbf_consts_struct(), flat.data_ptr(), node_in.data_ptr(),
                                                _lib.ptr(sign), pos.data_ptr(), plan.base.data_ptr(),
                                                plan.trip.data_ptr(), ws.data_ptr(), ws_bytes, int(need_grad),
                                                out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream,
                                                mod._aux_stream_ptr(dev), _lib.ptr(prepared)), "model_forward")
        if need_grad:
            ctx.mod, ctx.plan, ctx.ws, ctx.ws_bytes = mod, plan, ws, ws_bytes
            ctx.prepared = prepared
            ctx.node_in, ctx.sign, ctx.pos = node_in, sign, pos
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        mod, plan = ctx.mod, ctx.plan
        if ctx.ws is None:
            raise RuntimeError("pamnet_b200: backward through the same forward a second time (retain_graph=True) is not "
                               "supported -- the step's workspace is released after the first backward; run forward again")
        grad_out = grad_out.contiguous().float()
        target, direct = mod._grad_target()
        dev = grad_out.device
        with torch.cuda.device(dev):
            _lib.check(lib.pamnet_model_backward(mod._ccfg, plan.sizes, sbf_consts_struct(), mod._flat.data_ptr(),
                                                 ctx.node_in.data_ptr(), _lib.ptr(ctx.sign), ctx.pos.data_ptr(),
                                                 plan.base.data_ptr(), plan.trip.data_ptr(), ctx.ws.data_ptr(),
                                                 ctx.ws_bytes, grad_out.data_ptr(), target.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream,
                                                 mod._aux_stream_ptr(dev), _lib.ptr(ctx.prepared)),
                       "model_backward")
        ctx.ws = None
        mod._deliver_grads(target, direct)
        return (None, None, None, None, None, None, None)


class _PAMNetBase(nn.Module):
    """Shared machinery of PAMNet / PAMNet_s.

    Autograd contract (differs from a plain nn.Module): the module's parameters are views of one
    flat buffer and the whole model is ONE autograd node whose backward writes the gradients straight into the matching
    flat gradient buffer and attaches ``p.grad`` views.  Consequences:
      * ``loss.backward()`` (optionally after ``zero_grad``) is the supported way to obtain parameter gradients;
        ``torch.autograd.grad(loss, model.parameters())`` and ``backward(inputs=...)`` do not see the parameters, and
        parameter hooks do not fire -- so ``torch.nn.parallel.DistributedDataParallel`` would silently skip its
        all-reduce: wrapping is refused (use pamnet_b200.parallel instead);
      * a second backward through the same forward (retain_graph=True) raises;
      * grad tensors kept from an earlier step alias the flat buffer and are overwritten by the next backward."""
    _simple = False
    _MAX_NB = 1000         # max_num_neighbors of the radius graph: models.py:110,128 (PAMNet), :301 (PAMNet_s: 500)

    # ---- flat parameter storage ---------------------------------------------------------------------
    def _setup_flat(self, config):
        kind = _dataset_kind(config.dataset)
        flow = 1 if getattr(config, "flow", "source_to_target") == "target_to_source" else 0
        self._ccfg = _lib.Config(max(kind, 0), int(config.dim), int(config.n_layer), flow, int(self._simple),
                                 float(config.cutoff_l), float(config.cutoff_g))
        lib = _lib.load()
        n = lib.pamnet_param_count(self._ccfg)
        if n < 0:
            raise ValueError(lib.pamnet_last_error().decode())
        offs = (_lib.c_i64 * n)()
        numel = (_lib.c_i64 * n)()
        _lib.check(lib.pamnet_param_offsets(self._ccfg, offs, numel), "param_offsets")
        self._param_list = list(self.named_parameters())
        if len(self._param_list) != n or any(p.numel() != m for (_, p), m in zip(self._param_list, numel)):
            raise RuntimeError("module tree does not match the C parameter layout (state_dict order)")
        self._offsets = list(offs)
        self._total = int(lib.pamnet_param_total(self._ccfg))
        # init_linear exists but is unused outside PDBbind (models.py:35,119): its grad stays None there
        # and the atom-type embedding table is unused on PDBbind (models.py:119)
        self._param_used = [not ((name == "init_linear.weight" and kind != 1) or (name == "embeddings" and kind == 1))
                            for name, _ in self._param_list]
        self._flat = None
        self._gflat = None
        self._gviews = None
        self._alias_tick = 0
        self._flatten()

    def _aux_stream_ptr(self, dev):
        """Second CUDA stream for the x-independent GEMMs (PAMNET_STREAMS=1 disables the overlap)."""
        import os
        if os.environ.get("PAMNET_STREAMS", "2") == "1":
            return None
        aux = getattr(self, "_aux_stream", None)
        if aux is None or aux.device != dev:
            aux = self._aux_stream = torch.cuda.Stream(device=dev)
        return aux.cuda_stream

    # ---- gradients ------------------------------------------------------------------------------------
    def _grad_views(self):
        if self._gflat is None or self._gflat.device != self._flat.device:
            self._gflat = torch.zeros(self._total, dtype=torch.float32, device=self._flat.device)
            self._gviews = [self._gflat[off:off + p.numel()].view(p.shape)
                            for (_, p), off in zip(self._param_list, self._offsets)]
        return self._gviews

    def _views_attached(self):
        """O(1) check that the gradient views are (still) attached: first, middle and last parameter that receive one."""
        idx = getattr(self, "_sentinels", None)
        if idx is None:
            live = [i for i, ((_, p), used) in enumerate(zip(self._param_list, self._param_used)) if used and p.requires_grad]
            idx = self._sentinels = (live[0], live[len(live) // 2], live[-1]) if live else ()
        views = self._gviews
        return views is not None and all(self._param_list[i][1].grad is views[i] for i in idx)

    def zero_grad(self, set_to_none=True):
        """``optimizer.zero_grad()`` of main_qm9.py:106 for this module, without touching 390 tensors: when the gradients
        are the views backward attached, they STAY attached and the next backward overwrites the flat buffer (the per-step
        loops ``p.grad = None`` + re-attach cost ~0.2 ms of host time, and the step is host-bound).  Until that backward,
        ``p.grad`` shows the previous step's values instead of None; ``torch.optim.Optimizer.zero_grad`` (which sets None)
        keeps working and takes the ordinary path."""
        if set_to_none and self._gflat is not None and self._views_attached():
            self._overwrite_next = True
            return
        self._overwrite_next = False
        super().zero_grad(set_to_none)

    def _grad_target(self):
        """Where backward should write: straight into the flat gradient buffer when no parameter holds a
        gradient yet (after zero_grad), otherwise into a scratch buffer that is then added."""
        self._grad_views()
        if getattr(self, "_overwrite_next", False) and self._views_attached():
            return self._gflat, True
        self._overwrite_next = False
        if all(p.grad is None for _, p in self._param_list):
            return self._gflat, True
        return torch.empty_like(self._gflat), False

    def _deliver_grads(self, target, direct):
        views = self._grad_views()
        if direct and getattr(self, "_overwrite_next", False):
            self._overwrite_next = False          # the views were attached all along
            return
        if direct:
            for (_, p), v, used in zip(self._param_list, views, self._param_used):
                if used and p.requires_grad:
                    p.grad = v
            return
        attached = all((p.grad is v) or not (used and p.requires_grad)
                       for (_, p), v, used in zip(self._param_list, views, self._param_used))
        if attached:
            self._gflat.add_(target)           # one launch: every p.grad is a view of the flat buffer
            return
        for (_, p), off, used in zip(self._param_list, self._offsets, self._param_used):
            if not (used and p.requires_grad):
                continue
            g = target[off:off + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.add_(g)

    def _flatten(self):
        """(Re)pack all parameters into one flat buffer; parameters become views of it."""
        ps = [p for _, p in self._param_list]
        dev = ps[0].device
        if any(p.dtype != torch.float32 for p in ps):
            raise TypeError("pamnet_b200 computes in fp32 only (the reference's precision)")
        flat = torch.zeros(self._total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(ps, self._offsets):
                flat[off:off + p.numel()].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + p.numel()].view(p.shape)
        self._flat = flat.requires_grad_(True)     # the single differentiable input of _PAMNetFunction
        self._gflat = None
        self._gviews = None
        self._overwrite_next = False
        self._sentinels = None

    def _aliased(self, full=True):
        """Do the parameters still alias the flat buffer?  (utils/ema.py:27,32 swap param.data wholesale; a partial weight
        load may replace single tensors.)  All ~390 pointers are compared on every call (~55 us of host time; the round-1
        rotating window could miss a replaced tensor for up to 63 steps); `full` is kept for callers of the old signature.
        Replaced Parameter OBJECTS (load_state_dict(assign=True)) are picked up by load_state_dict below."""
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * off for (_, p), off in zip(self._param_list, self._offsets))

    def load_state_dict(self, state_dict, *args, **kwargs):
        out = super().load_state_dict(state_dict, *args, **kwargs)
        if getattr(self, "_param_list", None):
            cur = list(self.named_parameters())
            if any(a is not b for (_, a), (_, b) in zip(self._param_list, cur)):      # assign=True replaced the Parameters
                self._param_list = cur
                self._flatten()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_param_list", None):
            self._flatten()
        return out

    # ---- graph construction ---------------------------------------------------------------------------
    def _build_plan(self, pos, batch, n_graphs, edge_index_l_in, max_nb):
        """The front end in ONE C call (pamnet_plan_build): the two or three count read-backs it needs happen inside
        the library instead of through Python.  Buffers are allocated here at remembered capacities (the caller owns
        all memory); a too-small capacity makes the call return 1 with the needed sizes, and it is repeated once."""
        import os
        if os.environ.get("PAMNET_PLAN", "") == "stepwise":
            return self._build_plan_stepwise(pos, batch, n_graphs, edge_index_l_in, max_nb)
        lib = _lib.load()
        cfg, dev = self._ccfg, pos.device
        n, kind = pos.shape[0], cfg.dataset
        el_in = edge_index_l_in.to(torch.int64).contiguous() if kind == 0 else None
        e_in = int(el_in.shape[1]) if el_in is not None else 0
        caps = getattr(self, "_plan_caps", None)
        if caps is None:
            caps = {"eg": 50 * n if kind == 2 else 24 * n, "el": e_in if kind == 0 else (50 * n if kind == 2 else 8 * n),
                    "base": 0, "trip": 0}
        if kind == 0:
            caps["el"] = max(caps["el"], e_in)
        base_b, trip_b = _lib.c_sz(), _lib.c_sz()
        need = (_lib.c_i64 * 4)()
        sz = _lib.Sizes(n, n_graphs, 0, 0, 0, 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        for attempt in range(4):
            guess = _lib.Sizes(n, n_graphs, caps["eg"], caps["el"], 4 * caps["el"], 4 * caps["el"])
            _lib.check(lib.pamnet_plan_bytes(cfg, guess, base_b, trip_b), "plan_bytes")
            caps["base"], caps["trip"] = max(caps["base"], base_b.value), max(caps["trip"], trip_b.value)
            eg_buf = torch.empty(2 * max(caps["eg"], 1), dtype=torch.int64, device=dev)
            el_buf = torch.empty(2 * max(caps["el"], 1), dtype=torch.int64, device=dev)
            base = torch.empty(caps["base"], dtype=torch.uint8, device=dev)
            trip = torch.empty(caps["trip"], dtype=torch.uint8, device=dev)
            sb = lib.pamnet_plan_build_scratch_bytes(cfg, n, e_in, caps["eg"])
            scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                rc = lib.pamnet_plan_build(cfg, pos.data_ptr(), batch.data_ptr(), n, n_graphs, _lib.ptr(el_in), e_in,
                                           int(max_nb), eg_buf.data_ptr(), caps["eg"], el_buf.data_ptr(), caps["el"],
                                           base.data_ptr(), caps["base"], trip.data_ptr(), caps["trip"],
                                           scratch.data_ptr(), sb, sz, need, stream)
            if rc == 0:
                break
            if rc != 1:
                _lib.check(rc, "plan_build")
            grow = lambda have, want: max(have, int(want * 1.25) + 64)
            caps["eg"], caps["el"] = grow(caps["eg"], need[0]), grow(caps["el"], need[1])
            caps["base"], caps["trip"] = grow(caps["base"], need[2]), grow(caps["trip"], need[3])
        else:
            raise _lib.PamnetError("plan_build: capacities did not converge")
        self._plan_caps = caps
        e_g, e_l = int(sz.n_edges_g), int(sz.n_edges_l)
        eg = eg_buf[:2 * e_g].view(2, e_g)
        if kind == 0 and e_l == e_in:
            el = el_in
        elif kind == 1 and e_l == e_g:
            el = eg
        else:
            el = el_buf[:2 * e_l].view(2, e_l)
        return GraphPlan(sz, base, trip, eg, el)

    def _build_plan_stepwise(self, pos, batch, n_graphs, edge_index_l_in, max_nb):
        """Same plan through the granular operator calls (PAMNET_PLAN=stepwise; the path the operator tests cover)."""
        lib = _lib.load()
        cfg, dev = self._ccfg, pos.device
        n = pos.shape[0]
        c = ops.Counts(dev)
        kind = cfg.dataset
        if kind == 2:       # rna: kNN-50 then two cutoff masks (models.py:143-157)
            nbr, _ = ops.knn_lists(pos, batch, 50)
            pg = ops.knn_edges_count(nbr, pos, self.cutoff_g, c, 0)
            plc = ops.knn_edges_count(nbr, pos, self.cutoff_l, c, 1)
            eg_n, el_n = c.read()[:2]
            eg = ops.knn_edges_fill(nbr, pos, self.cutoff_g, pg, eg_n)
            el = ops.knn_edges_fill(nbr, pos, self.cutoff_l, plc, el_n)
        else:
            pg = ops.radius_count(pos, batch, self.cutoff_g, max_nb, True, c, 0)
            if kind == 0:   # QM9: local graph = chemical bonds (models.py:115)
                el_in = edge_index_l_in.to(torch.int64).contiguous()
                keep, pk = ops.edge_filter_count(el_in, None, None, c, 1)
                eg_n, el_n = c.read()[:2]
                eg = ops.radius_fill(pos, batch, self.cutoff_g, max_nb, True, pg, eg_n)
                el = ops.edge_filter_fill(el_in, keep, pk, el_n)
            else:           # PDBbind: local = global edges within cutoff_l (models.py:131-136)
                eg_n = c.read()[0]
                eg = ops.radius_fill(pos, batch, self.cutoff_g, max_nb, True, pg, eg_n)
                el = ops.filter_edges(eg, pos, self.cutoff_l)
        sz = _lib.Sizes(n, n_graphs, eg.shape[1], el.shape[1], 0, 0)
        base_b, trip_b = _lib.c_sz(), _lib.c_sz()
        _lib.check(lib.pamnet_plan_bytes(cfg, sz, base_b, trip_b), "plan_bytes")
        base = torch.empty(base_b.value, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.pamnet_plan_count(cfg, sz, eg.data_ptr(), el.data_ptr(), batch.data_ptr(), base.data_ptr(),
                                         c.ptr(2), stream), "plan_count")
        t2, t1 = c.read()[2:4]
        sz.n_t2, sz.n_t1 = t2, t1
        _lib.check(lib.pamnet_plan_bytes(cfg, sz, base_b, trip_b), "plan_bytes")
        trip = torch.empty(trip_b.value, dtype=torch.uint8, device=dev)
        _lib.check(lib.pamnet_plan_fill(cfg, sz, pos.data_ptr(), base.data_ptr(), trip.data_ptr(), stream), "plan_fill")
        return GraphPlan(sz, base, trip, eg, el)

    def _inputs(self, data):
        """The tensors the C calls read, in the layout they expect (models.py:101-106,117-125,138-141)."""
        x_raw, batch = data.x, data.batch
        if not x_raw.is_cuda:
            raise _lib.PamnetError("pamnet_b200 runs on CUDA tensors only: move the model and the batch to a GPU "
                                   "(there is no CPU fallback)")
        if self._flat.device != x_raw.device:
            raise RuntimeError("model and data are on different devices")
        kind = self._ccfg.dataset
        batch = batch.to(torch.int64).contiguous()
        n_graphs = getattr(data, "num_graphs", None)
        if n_graphs is None:
            n_graphs = int(batch.max()) + 1
        sign, el_in = None, None
        if kind == 0:
            pos = data.pos.to(torch.float32).contiguous()
            node_in = x_raw.to(torch.float32).contiguous().view(-1)
            el_in = data.edge_index
        else:
            xr = x_raw.unsqueeze(-1) if x_raw.dim() == 1 else x_raw
            xr = xr.to(torch.float32)
            pos = xr[:, :3].contiguous()
            if kind == 1:
                node_in = xr[:, 3:].contiguous()
                sign = torch.where(pos[:, 0] > 40.0, -1.0, 1.0).to(torch.float32).contiguous()   # models.py:122-125
            else:
                node_in = xr[:, -1].contiguous()
        return batch, int(n_graphs), pos, node_in, sign, el_in

    def prefetch(self, data, max_nb=None, wait_current=None):
        """Build the graph plan of ``data`` NOW on a side stream -- typically right after ``loss.backward()`` of the
        previous batch, so that the front end of models.py:104-177 (which does not depend on the parameters) overlaps the
        GPU work still queued for that step instead of preceding the first layer of the next one.  Returns the batch to
        call the model with: ``model(returned)`` on that SAME object starts from the attached plan.  The role a
        DataLoader worker plays for host-side preprocessing.  Opt-in; without it ``forward`` builds the plan itself.
        One batch can be pending at a time.

        ``data`` on the host (pinned): it is copied to the model's device on the side stream as well, so neither the
        copy nor the graph build waits for the backward queued on the current stream.  ``data`` already on the device:
        by default the side stream first waits for the current stream (the batch may have been produced there, and then
        nothing overlaps); pass ``wait_current=False`` when the batch is known to be complete."""
        dev = self._flat.device
        if dev.type != "cuda":
            raise _lib.PamnetError("pamnet_b200 runs on CUDA only: move the model to a GPU (there is no CPU fallback)")
        on_host = not data.x.is_cuda
        if wait_current is None:
            wait_current = not on_host
        cur = torch.cuda.current_stream(dev)
        side = getattr(self, "_side_stream", None)
        if side is None or side.device != dev:
            side = self._side_stream = torch.cuda.Stream(device=dev)
        if wait_current:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            if on_host:
                data = data.to(dev, non_blocking=True)
            inputs = self._inputs(data)
            batch, n_graphs, pos, node_in, sign, el_in = inputs
            plan = self._build_plan(pos, batch, n_graphs, el_in, self._MAX_NB if max_nb is None else max_nb)
            done = torch.cuda.Event()
            done.record(side)
        self._prefetched = (data, inputs, plan, done, on_host)
        return data

    def prefetch_async(self, data, max_nb=None, wait_current=None):
        """``prefetch`` on a worker thread: returns a handle at once; ``handle.result()`` is the batch to call the model
        with.  The plan build contains one small device-to-host read-back (the edge / triplet counts the buffer sizes
        depend on) and ~0.2 ms of host work; on the caller's thread that sits between ``backward()`` and the next
        ``forward()`` of a step whose host enqueue time equals its GPU time, on a worker it overlaps the enqueue of
        ``backward()`` (ctypes and torch release the GIL inside their calls).  Start it right after ``model(batch)``
        returned, collect it before the next ``model(...)``: what a ``DataLoader(num_workers=1)`` does for host-side
        preprocessing.  One request may be in flight at a time."""
        import concurrent.futures
        pool = getattr(self, "_prefetch_pool", None)
        if pool is None:
            pool = self._prefetch_pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="pamnet-prefetch")
        dev = self._flat.device
        cur = torch.cuda.current_stream(dev)
        on_host = not data.x.is_cuda
        if wait_current is None:
            wait_current = not on_host
        ready = None
        if wait_current:                      # "current stream" means the CALLER's: capture its state here, not on the worker
            ready = torch.cuda.Event()
            ready.record(cur)

        def work():
            with torch.cuda.device(dev):
                if ready is not None:
                    side = getattr(self, "_side_stream", None)
                    if side is None or side.device != dev:
                        side = self._side_stream = torch.cuda.Stream(device=dev)
                    side.wait_event(ready)
                return self.prefetch(data, max_nb=max_nb, wait_current=False)
        return pool.submit(work)

    def _take_prefetched(self, data):
        pre = getattr(self, "_prefetched", None)
        if pre is None or pre[0] is not data:
            return None
        self._prefetched = None
        _, inputs, plan, done, copied = pre
        cur = torch.cuda.current_stream(self._flat.device)
        cur.wait_event(done)
        # these blocks came from the side stream's pool and are consumed on the current stream from here on
        held = [*inputs, plan.base, plan.trip, plan.edge_index_g, plan.edge_index_l]
        if copied:          # the batch itself was allocated on the side stream; the caller goes on using it (data.y in the loss)
            fields = getattr(data, "__dict__", {})
            held += list(fields.values())
            store = fields.get("_store")
            if hasattr(store, "values"):
                held += list(store.values())
        for t in held:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(cur)
        return inputs, plan

    def _run(self, data, max_nb):
        if not self._aliased(full=False):      # e.g. EMA.assign swapped param.data (utils/ema.py:27)
            self._flatten()
        pre = self._take_prefetched(data)
        if pre is not None:
            (batch, n_graphs, pos, node_in, sign, el_in), plan = pre
            prepared = self._prepare_weights(pos.device)
        else:
            batch, n_graphs, pos, node_in, sign, el_in = self._inputs(data)
            prepared = self._prepare_weights(pos.device)      # on the auxiliary stream, overlapping the graph build
            plan = self._build_plan(pos, batch, n_graphs, el_in, max_nb)
        self.last_plan = plan
        return _PAMNetFunction.apply(self, plan, node_in, sign, pos, self._flat, prepared)

    def _prepare_weights(self, dev):
        """k-major chain weights + projection blocks (they depend on the parameters only) into a persistent blob,
        issued on the auxiliary stream so that they run while the host waits for the edge counts of the graph build.
        Returns None (model_forward then makes them itself) when there is no auxiliary stream."""
        import os
        aux_ptr = self._aux_stream_ptr(dev)
        # Off by default: measured from Python (B200, batch 32) the extra call + stream wait cost more host latency at the
        # start of the step (1.92 vs 1.89 ms/step) than the ~25 us of GPU time they take off the critical path.
        if aux_ptr is None or os.environ.get("PAMNET_PREP", "0") != "1":
            return None
        lib = _lib.load()
        buf = getattr(self, "_prepared", None)
        if buf is None or buf.device != dev:
            buf = self._prepared = torch.empty(lib.pamnet_prepared_weights_bytes(self._ccfg), dtype=torch.uint8, device=dev)
        self._aux_stream.wait_stream(torch.cuda.current_stream(dev))     # the parameters' last writer
        with torch.cuda.device(dev):
            _lib.check(lib.pamnet_prepare_weights(self._ccfg, self._flat.data_ptr(), buf.data_ptr(), aux_ptr), "prepare_weights")
        return buf

    def _init_embeddings(self):
        stdv = math.sqrt(3)
        self.embeddings.data.uniform_(-stdv, stdv)


class PAMNet(_PAMNetBase):
    def __init__(self, config: Config, num_spherical=7, num_radial=6, envelope_exponent=5):
        super(PAMNet, self).__init__()
        self.dataset = config.dataset
        self.dim = config.dim
        self.n_layer = config.n_layer
        self.cutoff_l = config.cutoff_l
        self.cutoff_g = config.cutoff_g
        rna = self.dataset[:3].lower() == "rna"
        self.embeddings = nn.Parameter(torch.ones((3 if rna else 5, self.dim)))
        if not rna:
            self.init_linear = nn.Linear(18, self.dim, bias=False)
        self.rbf_g = BesselBasisLayer(16, self.cutoff_g, envelope_exponent)
        self.rbf_l = BesselBasisLayer(16, self.cutoff_l, envelope_exponent)
        self.sbf = SphericalBasisLayer(num_spherical, num_radial, self.cutoff_l, envelope_exponent)
        self.mlp_rbf_g = MLP([16, self.dim])
        self.mlp_rbf_l = MLP([16, self.dim])
        self.mlp_sbf1 = MLP([num_spherical * num_radial, self.dim])
        self.mlp_sbf2 = MLP([num_spherical * num_radial, self.dim])
        self.global_layer = nn.ModuleList(Global_MessagePassing(config) for _ in range(config.n_layer))
        self.local_layer = nn.ModuleList(Local_MessagePassing(config) for _ in range(config.n_layer))
        self.softmax = nn.Softmax(dim=-1)
        self._init_embeddings()
        self._setup_flat(config)

    def forward(self, data):
        if _dataset_kind(self.dataset) < 0:
            raise ValueError("Invalid dataset. If you are using any dataset related to RNA 3D structure prediction, "
                             "be sure to use 'rna' as the first 3 characters of the dataset name.")
        return self._run(data, max_nb=self._MAX_NB)


class PAMNet_s(_PAMNetBase):
    _simple = True
    _MAX_NB = 500

    def __init__(self, config: Config, num_spherical=7, num_radial=6, envelope_exponent=5):
        super(PAMNet_s, self).__init__()
        self.dataset = config.dataset
        self.dim = config.dim
        self.n_layer = config.n_layer
        self.cutoff_l = config.cutoff_l
        self.cutoff_g = config.cutoff_g
        self.embeddings = nn.Parameter(torch.ones((5, self.dim)))
        self.rbf_g = BesselBasisLayer(16, self.cutoff_g, envelope_exponent)
        self.rbf_l = BesselBasisLayer(16, self.cutoff_l, envelope_exponent)
        self.sbf = SphericalBasisLayer(num_spherical, num_radial, self.cutoff_l, envelope_exponent)
        self.mlp_rbf_g = MLP([16, self.dim])
        self.mlp_rbf_l = MLP([16, self.dim])
        self.mlp_sbf = MLP([num_spherical * num_radial, self.dim])
        self.global_layer = nn.ModuleList(Global_MessagePassing(config) for _ in range(config.n_layer))
        self.local_layer = nn.ModuleList(Local_MessagePassing_s(config) for _ in range(config.n_layer))
        self.softmax = nn.Softmax(dim=-1)
        self._init_embeddings()
        if self.dataset == "QM9":
            self._setup_flat(config)

    def forward(self, data):
        if self.dataset != "QM9":
            raise ValueError("Invalid dataset. The current PAMNet_s is only for QM9 experiments.")
        return self._run(data, max_nb=self._MAX_NB)
