"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-ins for the four third-party packages the reference imports but which are
not installed anywhere in this image (torch_scatter 2.0.4, torch_sparse 0.6.0,
torch_cluster 1.5.4 via torch_geometric 1.4.2; pins: reference requirements.txt:8-11).
With these registered in ``sys.modules`` the reference's own ``models.py`` /
``layers/*.py`` / ``utils/sbf.py`` import and run UNMODIFIED from /root/reference,
which is how the restatement in ``oracle/pamnet_oracle.py`` is validated and how the
golden vectors in ``tests/golden`` were produced (``tests/golden/make_golden.py``).

The semantics below are OUR statement of what the pinned packages do (their source
is not in this container): **parity at these boundaries is unpinned by the
reference**, which ships no tests.

/root/reference only exists in the build container, never on the GPU box, so
nothing that runs under ``-m gpu``, ``smoke()`` or ``bench.py`` may call
``load_reference``.
"""
import inspect
import math
import os
import sys
import types

import numpy as np
import torch

REF_ROOT_DEFAULT = "/root/reference"


# --------------------------------------------------------------------------
# torch_scatter.scatter  (local_message_passing.py:50,54)
# --------------------------------------------------------------------------
def scatter(src, index, dim=0, out=None, dim_size=None, reduce="add"):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce in ("add", "sum"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        return res / cnt.view((-1,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


# --------------------------------------------------------------------------
# torch_sparse.SparseTensor -- only what models.py:72-96 touches
# --------------------------------------------------------------------------
class _Storage:
    def __init__(self, row, col, value):
        self._row, self._col, self._value = row, col, value

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value


class SparseTensor:
    def __init__(self, row, col, value=None, sparse_sizes=None, _sorted=False):
        n_rows, n_cols = sparse_sizes
        if not _sorted:
            key = row * n_cols + col
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            value = value[perm] if value is not None else None
        self._sizes = (n_rows, n_cols)
        self.storage = _Storage(row, col, value)
        counts = torch.bincount(row, minlength=n_rows)
        self._rowptr = torch.zeros(n_rows + 1, dtype=torch.long, device=row.device)
        self._rowptr[1:] = torch.cumsum(counts, 0)

    def __getitem__(self, idx):
        # row gather: new row r holds old row idx[r], column order preserved
        assert isinstance(idx, torch.Tensor) and idx.dtype == torch.long and idx.dim() == 1
        start = self._rowptr[idx]
        cnt = self._rowptr[idx + 1] - start
        total = int(cnt.sum())
        new_row = torch.repeat_interleave(torch.arange(idx.numel(), device=idx.device), cnt)
        first = torch.cumsum(cnt, 0) - cnt
        within = torch.arange(total, device=idx.device) - first[new_row]
        src = start[new_row] + within
        val = self.storage._value[src] if self.storage._value is not None else None
        return SparseTensor(new_row, self.storage._col[src], val,
                            sparse_sizes=(idx.numel(), self._sizes[1]), _sorted=True)

    def set_value(self, value, layout=None):
        return SparseTensor(self.storage._row, self.storage._col, value,
                            sparse_sizes=self._sizes, _sorted=True)

    def sum(self, dim):
        assert dim == 1
        if self.storage._value is None:
            return (self._rowptr[1:] - self._rowptr[:-1])
        out = torch.zeros(self._sizes[0], dtype=self.storage._value.dtype)
        return out.index_add_(0, self.storage._row, self.storage._value)


# --------------------------------------------------------------------------
# torch_cluster.radius / knn (re-exported by torch_geometric.nn; models.py:6,110,143)
# Canonical semantics: see oracle/graph_ops.py (kept in ONE place).
# --------------------------------------------------------------------------
def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    from oracle.graph_ops import radius_pairs
    return radius_pairs(x, y, r, batch_x, batch_y, max_num_neighbors)


def knn(x, y, k, batch_x=None, batch_y=None):
    from oracle.graph_ops import knn_pairs
    return knn_pairs(x, y, k, batch_x, batch_y)


# --------------------------------------------------------------------------
# torch_geometric bits
# --------------------------------------------------------------------------
def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="add")


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="mean")


class MessagePassing(torch.nn.Module):
    """aggr='add' message passing with the _i/_j argument convention of PyG 1.4."""

    def __init__(self, aggr="add", flow="source_to_target"):
        super().__init__()
        assert aggr == "add" and flow in ("source_to_target", "target_to_source")
        self.flow = flow
        self._msg_args = list(inspect.signature(self.message).parameters)

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (0, 1) if self.flow == "target_to_source" else (1, 0)
        kwargs = dict(kwargs)
        kwargs["edge_index"] = edge_index
        n_out = None
        args = []
        for name in self._msg_args:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                sel = edge_index[i] if name.endswith("_i") else edge_index[j]
                args.append(base.index_select(0, sel))
                n_out = base.size(0)
            else:
                args.append(kwargs[name])
        out = self.message(*args)
        out = scatter(out, edge_index[i], dim=0, dim_size=n_out, reduce="add")
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out


class Data:
    """Duck-typed torch_geometric Batch: just attributes plus ``.to``."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


def install():
    """Register the stand-ins; idempotent."""
    if not hasattr(np, "math"):
        np.math = math  # utils/sbf.py:65 uses np.math.factorial (removed in numpy 2)
    if "torch_scatter" in sys.modules and getattr(sys.modules["torch_scatter"], "_pamnet_shim", False):
        return
    ts = types.ModuleType("torch_scatter")
    ts.scatter = scatter
    ts._pamnet_shim = True
    tsp = types.ModuleType("torch_sparse")
    tsp.SparseTensor = SparseTensor
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.MessagePassing = MessagePassing
    tgnn.global_add_pool = global_add_pool
    tgnn.global_mean_pool = global_mean_pool
    tgnn.radius = radius
    tgnn.knn = knn
    tginits = types.ModuleType("torch_geometric.nn.inits")
    tginits.glorot = glorot
    tgutils = types.ModuleType("torch_geometric.utils")
    tgutils.remove_self_loops = remove_self_loops
    tgdata = types.ModuleType("torch_geometric.data")
    tgdata.Data = Data
    tg.nn, tg.utils, tg.data = tgnn, tgutils, tgdata
    tgnn.inits = tginits
    sys.modules.update({
        "torch_scatter": ts, "torch_sparse": tsp, "torch_geometric": tg,
        "torch_geometric.nn": tgnn, "torch_geometric.nn.inits": tginits,
        "torch_geometric.utils": tgutils, "torch_geometric.data": tgdata,
    })


def reference_available(ref_root=None):
    ref_root = ref_root or os.environ.get("PAMNET_REFERENCE_ROOT", REF_ROOT_DEFAULT)
    return os.path.isfile(os.path.join(ref_root, "models.py"))


def load_reference(ref_root=None):
    """Import the reference's models.py verbatim; returns the module.

    The reference uses top-level package names ``layers`` / ``utils`` / ``models``; they are
    imported under those names with ref_root temporarily first on sys.path.
    """
    ref_root = ref_root or os.environ.get("PAMNET_REFERENCE_ROOT", REF_ROOT_DEFAULT)
    if not reference_available(ref_root):
        raise FileNotFoundError(f"reference not found under {ref_root}")
    install()
    clash = [m for m in ("models", "layers", "utils") if m in sys.modules
             and not getattr(sys.modules[m], "__file__", "").startswith(ref_root)]
    for m in clash:
        del sys.modules[m]
    sys.path.insert(0, ref_root)
    try:
        import models  # noqa: the reference's module
    finally:
        sys.path.remove(ref_root)
    return models
