"""Operator surface: the third-party calls the reference model makes, on hand-written sm_100a kernels.

reference call (file:line)                                     -> here
torch_cluster.radius            models.py:110,128,301          -> radius / radius_graph
torch_cluster.knn               models.py:143                  -> knn
remove_self_loops               models.py:63                   -> remove_self_loops
dist mask                       models.py:131-136,148-157      -> filter_edges / knn_edges
PAMNet.indices (torch_sparse)   models.py:68-98                -> triplet_indices
torch_scatter.scatter           local_message_passing.py:50,54 -> scatter
BesselBasisLayer.forward        layers/basic.py:74-76          -> bessel_rbf
SphericalBasisLayer.forward     layers/basic.py:107-116        -> spherical_basis
nn.Linear (+SiLU)               layers/basic.py:19-22          -> linear
F.l1_loss / MSE                 main_qm9.py:108, main_pdbbind.py -> l1_loss / mse_loss (loss + gradient in one launch)

All tensors must live on a CUDA device; there is no CPU path.
"""
import torch

from . import _lib
from .basis import sbf_consts_struct


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    """CUDA tensors on the CURRENT device: the operator calls enqueue on torch's current stream and the library launches on
    the current device, so a tensor of another GPU would be addressed from the wrong device."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.PamnetError("pamnet_b200 ops run on CUDA tensors only (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.PamnetError(f"pamnet_b200 ops: tensor on cuda:{t.device.index} but cuda:{cur} is the current device; "
                                   f"call inside `with torch.cuda.device({t.device.index}):`")


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _i64(t):
    return t.detach().to(torch.int64).contiguous()


class Counts:
    """A few device int64 counters read back with one synchronising copy."""

    def __init__(self, device, n=8):
        self.dev = torch.zeros(n, dtype=torch.int64, device=device)

    def ptr(self, i):
        return self.dev.data_ptr() + 8 * i

    def read(self):
        return self.dev.cpu().tolist()


# ------------------------------------------------------------------------------------------------
def radius_count(pos, batch, r, max_num_neighbors, drop_self, counts, slot):
    lib = _lib.load()
    n = pos.shape[0]
    deg = torch.empty(max(n, 1), dtype=torch.int32, device=pos.device)
    ptr = torch.empty(n + 1, dtype=torch.int32, device=pos.device)
    _lib.check(lib.pamnet_radius_count(pos.data_ptr(), batch.data_ptr(), n, float(r), int(max_num_neighbors),
                                       int(drop_self), deg.data_ptr(), ptr.data_ptr(), counts.ptr(slot), _stream()),
               "radius_count")
    return ptr


def radius_fill(pos, batch, r, max_num_neighbors, drop_self, ptr, total):
    lib = _lib.load()
    ei = torch.empty((2, total), dtype=torch.int64, device=pos.device)
    _lib.check(lib.pamnet_radius_fill(pos.data_ptr(), batch.data_ptr(), pos.shape[0], float(r),
                                      int(max_num_neighbors), int(drop_self), ptr.data_ptr(), total,
                                      ei.data_ptr(), _stream()), "radius_fill")
    return ei


def radius_graph_grid(pos, batch, r, max_num_neighbors=32, loop=False, n_graphs=None):
    """radius_graph through the cell-list kernels (csrc/graph_grid.cu): the same edge list, bit for bit; meant for
    graphs of hundreds to thousands of atoms (PDBbind complexes, models.py:128)."""
    _require_cuda(pos, batch)
    lib = _lib.load()
    pos, batch = _f32(pos), _i64(batch)
    n = pos.shape[0]
    if n_graphs is None:
        n_graphs = int(batch.max()) + 1 if n else 1
    c = Counts(pos.device)
    sb = lib.pamnet_radius_grid_scratch_bytes(n, n_graphs)
    scratch = torch.empty(sb, dtype=torch.uint8, device=pos.device)
    deg = torch.empty(max(n, 1), dtype=torch.int32, device=pos.device)
    ptr = torch.empty(n + 1, dtype=torch.int32, device=pos.device)
    _lib.check(lib.pamnet_radius_grid_count(pos.data_ptr(), batch.data_ptr(), n, n_graphs, float(r), int(max_num_neighbors),
                                            int(not loop), scratch.data_ptr(), sb, deg.data_ptr(), ptr.data_ptr(), c.ptr(0),
                                            _stream()), "radius_grid_count")
    total = c.read()[0]
    ei = torch.empty((2, total), dtype=torch.int64, device=pos.device)
    _lib.check(lib.pamnet_radius_grid_fill(pos.data_ptr(), batch.data_ptr(), n, n_graphs, float(r), int(max_num_neighbors),
                                           int(not loop), scratch.data_ptr(), ptr.data_ptr(), total, ei.data_ptr(), _stream()),
               "radius_grid_fill")
    return ei


def radius_graph(pos, batch, r, max_num_neighbors=32, loop=False, method="auto"):
    """edge_index [2,E] (row = query, col = neighbour), ordered by (query, neighbour).  method: "brute" (per-graph scan,
    molecules), "grid" (cell list, large graphs) or "auto" (by atoms per graph)."""
    _require_cuda(pos, batch)
    pos, batch = _f32(pos), _i64(batch)
    if method == "grid" or (method == "auto" and pos.shape[0] >= 4096):      # auto: only worth a batch.max() sync when large
        n_graphs = int(batch.max()) + 1 if pos.shape[0] else 1
        if method == "grid" or pos.shape[0] // n_graphs >= 192:
            return radius_graph_grid(pos, batch, r, max_num_neighbors, loop, n_graphs)
    c = Counts(pos.device)
    ptr = radius_count(pos, batch, r, max_num_neighbors, not loop, c, 0)
    return radius_fill(pos, batch, r, max_num_neighbors, not loop, ptr, c.read()[0])


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """torch_cluster.radius for the self-query form the reference uses (x is y); returns (row, col)."""
    if x is not y and not (x.shape == y.shape and torch.equal(x, y)):
        raise NotImplementedError("only radius(pos, pos, ...) is on the hot path (models.py:110)")
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
    ei = radius_graph(x, batch_x, r, max_num_neighbors, loop=True)
    return ei[0], ei[1]


def knn_lists(pos, batch, k):
    _require_cuda(pos, batch)
    lib = _lib.load()
    pos, batch = _f32(pos), _i64(batch)
    n = pos.shape[0]
    nbr = torch.empty((n, k), dtype=torch.int32, device=pos.device)
    d2 = torch.empty((n, k), dtype=torch.float32, device=pos.device)
    _lib.check(lib.pamnet_knn(pos.data_ptr(), batch.data_ptr(), n, k, nbr.data_ptr(), d2.data_ptr(), _stream()), "knn")
    return nbr, d2


def knn(x, y, k, batch_x=None, batch_y=None):
    """torch_cluster.knn(pos, pos, k, batch, batch) -> (row = query, col = neighbour), ascending distance."""
    if x is not y and not (x.shape == y.shape and torch.equal(x, y)):
        raise NotImplementedError("only knn(pos, pos, ...) is on the hot path (models.py:143)")
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
    nbr, _ = knn_lists(x, batch_x, k)
    n = nbr.shape[0]
    row = torch.arange(n, device=x.device).repeat_interleave(k)
    col = nbr.reshape(-1).to(torch.int64)
    keep = col >= 0
    return row[keep], col[keep]


def knn_edges_count(nbr, pos, cutoff, counts, slot):
    lib = _lib.load()
    n, k = nbr.shape
    deg = torch.empty(max(n, 1), dtype=torch.int32, device=pos.device)
    ptr = torch.empty(n + 1, dtype=torch.int32, device=pos.device)
    _lib.check(lib.pamnet_knn_edges_count(nbr.data_ptr(), pos.data_ptr(), n, k, float(cutoff), deg.data_ptr(),
                                          ptr.data_ptr(), counts.ptr(slot), _stream()), "knn_edges_count")
    return ptr


def knn_edges_fill(nbr, pos, cutoff, ptr, total):
    lib = _lib.load()
    n, k = nbr.shape
    ei = torch.empty((2, total), dtype=torch.int64, device=pos.device)
    _lib.check(lib.pamnet_knn_edges_fill(nbr.data_ptr(), pos.data_ptr(), n, k, float(cutoff), ptr.data_ptr(), total,
                                         ei.data_ptr(), _stream()), "knn_edges_fill")
    return ei


def edge_filter_count(edge_index, pos, cutoff, counts, slot):
    lib = _lib.load()
    e = edge_index.shape[1]
    keep = torch.empty(max(e, 1), dtype=torch.int32, device=edge_index.device)
    ptr = torch.empty(e + 1, dtype=torch.int32, device=edge_index.device)
    _lib.check(lib.pamnet_edge_filter_count(edge_index.data_ptr(), e, _lib.ptr(pos),
                                            float(cutoff if cutoff is not None else 0.0), keep.data_ptr(),
                                            ptr.data_ptr(), counts.ptr(slot), _stream()), "edge_filter_count")
    return keep, ptr


def edge_filter_fill(edge_index, keep, ptr, total):
    lib = _lib.load()
    e = edge_index.shape[1]
    if total == e:
        return edge_index
    out = torch.empty((2, total), dtype=torch.int64, device=edge_index.device)
    _lib.check(lib.pamnet_edge_filter_fill(edge_index.data_ptr(), e, keep.data_ptr(), ptr.data_ptr(), total,
                                           out.data_ptr(), _stream()), "edge_filter_fill")
    return out


def filter_edges(edge_index, pos=None, cutoff=None):
    """Order-preserving: drop self loops and (optionally) edges longer than ``cutoff``."""
    _require_cuda(edge_index, pos)
    edge_index = _i64(edge_index)
    pos = _f32(pos) if pos is not None else None
    c = Counts(edge_index.device)
    keep, ptr = edge_filter_count(edge_index, pos, cutoff, c, 0)
    return edge_filter_fill(edge_index, keep, ptr, c.read()[0])


def remove_self_loops(edge_index, edge_attr=None):
    if edge_attr is not None:
        raise NotImplementedError("edge_attr is never passed on the hot path (models.py:63)")
    return filter_edges(edge_index), None


TRIPLET_NAMES = ("idx_i", "idx_j", "idx_k", "idx_kj", "idx_ji",
                 "idx_i_pair", "idx_j1_pair", "idx_j2_pair", "idx_jj_pair", "idx_ji_pair")


def triplet_indices(edge_index, num_nodes):
    """PAMNet.indices (models.py:68-98): the ten int64 index vectors, reference order."""
    _require_cuda(edge_index)
    lib = _lib.load()
    edge_index = _i64(edge_index)
    e, dev = edge_index.shape[1], edge_index.device
    scratch = torch.empty(lib.pamnet_triplet_scratch_bytes(num_nodes, e), dtype=torch.uint8, device=dev)
    c = Counts(dev)
    _lib.check(lib.pamnet_triplet_count(edge_index.data_ptr(), e, num_nodes, scratch.data_ptr(), c.ptr(0), _stream()),
               "triplet_count")
    t2, t1 = c.read()[:2]
    outs = [torch.empty(t2 if i < 5 else t1, dtype=torch.int64, device=dev) for i in range(10)]
    _lib.check(lib.pamnet_triplet_fill(edge_index.data_ptr(), e, num_nodes, scratch.data_ptr(), t2, t1,
                                       *[o.data_ptr() for o in outs], _stream()), "triplet_fill")
    return tuple(outs)


def _scatter_add(src, index, dim_size):
    lib = _lib.load()
    width = 1
    for d in src.shape[1:]:
        width *= d
    src2 = _f32(src).reshape(src.shape[0], width)
    out = torch.empty((dim_size, src2.shape[1]), dtype=torch.float32, device=src.device)
    _lib.check(lib.pamnet_scatter_add(src2.data_ptr(), index.data_ptr(), src2.shape[0], src2.shape[1], dim_size,
                                      out.data_ptr(), _stream()), "scatter_add")
    return out.reshape((dim_size,) + tuple(src.shape[1:]))


class _ScatterFn(torch.autograd.Function):
    """out[index[k]] += src[k]; the gradient of a segment sum is a row gather."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        ctx.save_for_backward(index)
        return _scatter_add(src, index, dim_size)

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        return g.contiguous().index_select(0, index), None, None


def scatter(src, index, dim=0, dim_size=None, reduce="add"):
    """torch_scatter.scatter(src, index, dim=0, dim_size=n, reduce='add') as the hot path calls it
    (local_message_passing.py:50,54); differentiable w.r.t. ``src``."""
    if dim != 0 or reduce not in ("add", "sum"):
        raise NotImplementedError("hot path uses scatter(src, index, dim=0, reduce='add') only")
    _require_cuda(src, index)
    index = _i64(index)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    if torch.is_grad_enabled() and src.requires_grad:
        return _ScatterFn.apply(src.float(), index, dim_size)
    return _scatter_add(src, index, dim_size)


def bessel_rbf(dist, freq, cutoff):
    _require_cuda(dist, freq)
    lib = _lib.load()
    dist, freq = _f32(dist), _f32(freq)
    out = torch.empty((dist.shape[0], 16), dtype=torch.float32, device=dist.device)
    _lib.check(lib.pamnet_bessel_rbf(dist.data_ptr(), dist.shape[0], freq.data_ptr(), float(cutoff), out.data_ptr(),
                                     _stream()), "bessel_rbf")
    return out


def spherical_basis(dist, angle, idx_kj, cutoff):
    _require_cuda(dist, angle, idx_kj)
    lib = _lib.load()
    dist, angle, idx_kj = _f32(dist), _f32(angle), _i64(idx_kj)
    consts = sbf_consts_struct()
    radial = torch.empty((dist.shape[0], 42), dtype=torch.float32, device=dist.device)
    _lib.check(lib.pamnet_sbf_radial(consts, dist.data_ptr(), dist.shape[0], float(cutoff), radial.data_ptr(),
                                     _stream()), "sbf_radial")
    out = torch.empty((angle.shape[0], 42), dtype=torch.float32, device=dist.device)
    _lib.check(lib.pamnet_spherical_basis(consts, radial.data_ptr(), angle.data_ptr(), idx_kj.data_ptr(),
                                          angle.shape[0], out.data_ptr(), _stream()), "spherical_basis")
    return out


def _linear_fwd(x, weight, bias, silu):
    lib = _lib.load()
    y = torch.empty((x.shape[0], weight.shape[0]), dtype=torch.float32, device=x.device)
    _lib.check(lib.pamnet_linear(x.data_ptr(), x.shape[0], weight.shape[1], weight.shape[0], weight.data_ptr(),
                                 _lib.ptr(bias), int(silu), y.data_ptr(), _stream()), "linear")
    return y


class _LinearFn(torch.autograd.Function):
    """y = [SiLU](x W^T + b) (layers/basic.py:19-22) with the hand-written GEMM kernels in both directions:
    g_x = g_z W (pamnet_gemm mode 1), g_W = g_z^T x and g_b = column sums of g_z (mode 2, split over the rows)."""

    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        x, weight = x.contiguous(), weight.contiguous()
        bias = bias.contiguous() if bias is not None else None
        z = _linear_fwd(x, weight, bias, False)
        ctx.save_for_backward(x, weight, z if silu else None)
        ctx.silu, ctx.has_bias = silu, bias is not None
        return z * torch.sigmoid(z) if silu else z

    @staticmethod
    def backward(ctx, g):
        x, weight, z = ctx.saved_tensors
        g = g.contiguous().float()
        if ctx.silu:                       # SiLU'(z) = s (1 + z (1 - s))
            s = torch.sigmoid(z)
            g = g * (s * (1.0 + z * (1.0 - s)))
        rows, n_in, n_out = x.shape[0], weight.shape[1], weight.shape[0]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm(1, g, weight, rows, n_in, n_out)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = gemm(2, g, x, n_out, n_in, rows, ksplit=max(1, min(64, (rows + 127) // 128)), want_dbias=True)
        return gx, gw, (gb if ctx.has_bias else None), None


def linear(x, weight, bias=None, silu=False):
    """nn.Linear (+ SiLU) on the CUDA GEMM kernels; differentiable w.r.t. x, weight and bias."""
    _require_cuda(x, weight, bias)
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        return _LinearFn.apply(x.float(), weight.float(), bias.float() if bias is not None else None, bool(silu))
    return _linear_fwd(_f32(x), _f32(weight), _f32(bias) if bias is not None else None, silu)


def gemm(mode, a, b, m, n, k, ksplit=1, want_dbias=False):
    """Test hook for the fp32 GEMM paths (mode 0: A B^T, 1: A B, 2: A^T B)."""
    _require_cuda(a, b)
    lib = _lib.load()
    a, b = _f32(a), _f32(b)
    c = torch.zeros((m, n), dtype=torch.float32, device=a.device)
    dbias = torch.zeros(m, dtype=torch.float32, device=a.device) if want_dbias else None
    _lib.check(lib.pamnet_gemm(mode, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], c.data_ptr(), n, m, n, k,
                               ksplit, _lib.ptr(dbias), _stream()), "gemm")
    return (c, dbias) if want_dbias else c


class _FusedLoss(torch.autograd.Function):
    """mean |out - y| (kind 0, F.l1_loss of main_qm9.py:108) or mean (out - y)^2 (kind 1, the MSE of main_pdbbind.py) with
    its gradient in ONE launch (pamnet_loss); torch's own sequence is ~7 tiny launches between forward and backward."""

    @staticmethod
    def forward(ctx, out, y, kind):
        _require_cuda(out, y)
        if out.shape != y.shape:
            raise ValueError(f"loss: prediction {tuple(out.shape)} and target {tuple(y.shape)} differ in shape")
        o, t = _f32(out).reshape(-1), _f32(y).reshape(-1)
        loss = torch.empty(1, dtype=torch.float32, device=o.device)
        grad = torch.empty_like(o)
        _lib.check(_lib.load().pamnet_loss(o.data_ptr(), t.data_ptr(), o.numel(), int(kind), loss.data_ptr(),
                                           grad.data_ptr(), _stream()), "loss")
        ctx.save_for_backward(grad)
        ctx.shape = out.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None, None


def l1_loss(out, y):
    """F.l1_loss(out, y) (mean reduction) with gradient w.r.t. ``out`` only (the target is data)."""
    return _FusedLoss.apply(out, y, 0)


def mse_loss(out, y):
    """F.mse_loss(out, y) (mean reduction) with gradient w.r.t. ``out`` only."""
    return _FusedLoss.apply(out, y, 1)
