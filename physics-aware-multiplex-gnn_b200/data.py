"""Synthetic QM9-shaped batches and a duck-typed ``Batch``.

Real QM9 cannot be fetched here (no network, no rdkit; reference datasets/qm9_dataset.py:116-168),
so every benchmark and parity test uses this generator.  The field layout is the one the
reference model reads (models.py:101-106; datasets/qm9_dataset.py:231-251):

* ``x``          [N]    float32 atom-type id in 0..4
* ``pos``        [N,3]  float32
* ``edge_index`` [2,E]  int64, chemical bonds in both directions, per molecule sorted by
                        row*n+col, offset by the collate
* ``batch``      [N]    int64 non-decreasing graph id
* ``y``          [G]    float32 target

RNA-shaped batches follow datasets/tu_dataset.py:111-115: ``x`` [N,4] = (xyz, type in 0..2),
no ``pos`` / ``edge_index``.
"""
import numpy as np
import torch


class Batch:
    """Minimal stand-in for torch_geometric.data.Batch: attributes, ``to``, ``num_graphs``."""

    def __init__(self, **fields):
        self.__dict__.update(fields)

    @property
    def num_graphs(self):
        return int(self.y.shape[0]) if getattr(self, "y", None) is not None else int(self.batch.max()) + 1

    def to(self, device, non_blocking=False):
        packed = self.__dict__.get("_packed")
        if packed is not None and torch.device(device).type == "cuda":
            # one host->device copy of the packed pinned blob, fields are views of the device blob
            blob, table = packed
            dev = blob.to(device, non_blocking=non_blocking)
            out = Batch()
            for k, v in self.__dict__.items():
                if k != "_packed" and not isinstance(v, torch.Tensor):
                    out.__dict__[k] = v
            for k, off, nbytes, dtype, shape in table:
                out.__dict__[k] = dev[off:off + nbytes].view(dtype).view(shape)
            return out
        out = Batch()
        for k, v in self.__dict__.items():
            if k == "_packed":
                continue
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        return out

    def pin_memory(self):
        """All tensor fields packed into ONE pinned blob (256 B aligned segments): ``.to('cuda')`` is then a single
        cudaMemcpyAsync instead of one per field (the reference's ``data.to(device)``, main_qm9.py:104, issues five)."""
        tensors = [(k, v.contiguous()) for k, v in self.__dict__.items() if isinstance(v, torch.Tensor)]
        table, off = [], 0
        for k, v in tensors:
            nbytes = v.numel() * v.element_size()
            table.append((k, off, nbytes, v.dtype, tuple(v.shape)))
            off += (nbytes + 255) // 256 * 256
        blob = torch.empty(max(off, 1), dtype=torch.uint8).pin_memory()
        out = Batch()
        for k, v in self.__dict__.items():
            if not isinstance(v, torch.Tensor) and k != "_packed":
                out.__dict__[k] = v
        for (k, v), (_, o, nbytes, dtype, shape) in zip(tensors, table):
            view = blob[o:o + nbytes].view(dtype).view(shape)
            view.copy_(v)
            out.__dict__[k] = view
        out.__dict__["_packed"] = (blob, table)
        return out


def _grow_molecule(rng, n_atoms, min_sep=0.9, ring_cut=1.7, max_deg=4):
    pos = np.zeros((n_atoms, 3))
    deg = np.zeros(n_atoms, dtype=np.int64)
    bonds = set()
    k = 1
    while k < n_atoms:
        cand = np.flatnonzero(deg[:k] < max_deg)
        parent = int(rng.choice(cand))
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        p = pos[parent] + v * rng.uniform(1.0, 1.6)
        if np.min(np.linalg.norm(pos[:k] - p, axis=1)) < min_sep:
            continue
        pos[k] = p
        bonds.add((parent, k))
        deg[parent] += 1
        deg[k] += 1
        k += 1
    for a in range(n_atoms):                      # ring closures
        for b in range(a + 1, n_atoms):
            if (a, b) in bonds or deg[a] >= max_deg or deg[b] >= max_deg:
                continue
            if np.linalg.norm(pos[a] - pos[b]) < ring_cut:
                bonds.add((a, b))
                deg[a] += 1
                deg[b] += 1
    return pos, sorted(bonds)


def synthetic_qm9_batch(num_graphs=32, seed=0, min_atoms=12, max_atoms=28):
    """Random-growth molecules: bond length U(1.0,1.6), min separation 0.9, degree <= 4."""
    rng = np.random.default_rng(seed)
    xs, poss, rows, cols, batch = [], [], [], [], []
    offset = 0
    for g in range(num_graphs):
        n = int(rng.integers(min_atoms, max_atoms + 1))
        pos, bonds = _grow_molecule(rng, n)
        r = np.array([a for a, b in bonds] + [b for a, b in bonds], dtype=np.int64)
        c = np.array([b for a, b in bonds] + [a for a, b in bonds], dtype=np.int64)
        order = np.argsort(r * n + c, kind="stable")
        rows.append(r[order] + offset)
        cols.append(c[order] + offset)
        xs.append(rng.integers(0, 5, size=n).astype(np.float32))
        poss.append(pos.astype(np.float32))
        batch.append(np.full(n, g, dtype=np.int64))
        offset += n
    y = rng.normal(size=num_graphs).astype(np.float32)
    return Batch(
        x=torch.from_numpy(np.concatenate(xs)),
        pos=torch.from_numpy(np.concatenate(poss)),
        edge_index=torch.from_numpy(np.stack([np.concatenate(rows), np.concatenate(cols)])),
        batch=torch.from_numpy(np.concatenate(batch)),
        y=torch.from_numpy(y),
    )


def synthetic_rna_batch(num_graphs=2, seed=0, min_atoms=200, max_atoms=400, spacing=1.5):
    """RNA-like point clouds: a self-avoiding-ish chain with side atoms, 3-decimal coordinates,
    atom types {0,1,2}; x = [pos, type] (datasets/tu_dataset.py:111-115)."""
    rng = np.random.default_rng(seed)
    xs, batch = [], []
    for g in range(num_graphs):
        n = int(rng.integers(min_atoms, max_atoms + 1))
        pos = np.zeros((n, 3))
        direction = rng.normal(size=3)
        direction /= np.linalg.norm(direction)
        k = 1
        while k < n:
            direction = direction + 0.6 * rng.normal(size=3)
            direction /= np.linalg.norm(direction)
            anchor = pos[k - 1] if k % 3 else pos[max(k - 3, 0)]
            p = anchor + direction * rng.uniform(0.9 * spacing, 1.1 * spacing)
            if np.min(np.linalg.norm(pos[:k] - p, axis=1)) < 1.1:
                continue
            pos[k] = p
            k += 1
        pos = np.round(pos + rng.uniform(50, 200, size=3), 3)
        t = rng.integers(0, 3, size=n).astype(np.float64)
        xs.append(np.concatenate([pos, t[:, None]], 1).astype(np.float32))
        batch.append(np.full(n, g, dtype=np.int64))
    y = rng.uniform(0, 10, size=num_graphs).astype(np.float32)
    return Batch(x=torch.from_numpy(np.concatenate(xs)),
                 batch=torch.from_numpy(np.concatenate(batch)),
                 y=torch.from_numpy(y))


def molecules_of(batch):
    """Split a collated QM9-shaped ``Batch`` back into per-molecule records (x [n], pos [n,3], edge_index [2,e] with atom
    ids inside the molecule, y scalar) -- the form a dataset's ``Data`` objects have before the loader collates them."""
    b = batch.batch.cpu().numpy()
    x, pos = batch.x.cpu().numpy(), batch.pos.cpu().numpy()
    ei, y = batch.edge_index.cpu().numpy(), batch.y.cpu().numpy()
    n_graphs = int(y.shape[0])
    starts = np.searchsorted(b, np.arange(n_graphs + 1), side="left")
    eg = b[ei[0]] if ei.shape[1] else np.zeros(0, dtype=np.int64)
    out = []
    for g in range(n_graphs):
        s, e = int(starts[g]), int(starts[g + 1])
        sel = np.flatnonzero(eg == g)
        out.append(types_ns(x=x[s:e].copy(), pos=pos[s:e].copy(), edge_index=ei[:, sel] - s, y=float(y[g])))
    return out


def types_ns(**kw):
    import types
    return types.SimpleNamespace(**kw)


class DeviceDataset:
    """QM9-shaped molecules resident in device memory; ``batch(ids)`` collates on the device in ONE kernel launch
    (pamnet_collate) -- SURVEY.md 8(f) row 2.  Replaces the host-side collate of PyG's DataLoader plus ``data.to(device)``
    (main_qm9.py:59-60,103-104): per batch the host sends only a [3, G] int64 table (molecule id, first atom and first
    bond of the molecule inside the batch) instead of every field of every atom.

    ``molecules``: sequence of records with ``x`` [n], ``pos`` [n, 3], ``edge_index`` [2, e] (atom ids inside the
    molecule) and scalar ``y`` -- torch tensors or numpy arrays.  CUDA only: there is no CPU path."""

    def __init__(self, molecules, device):
        device = torch.device(device)
        if device.type != "cuda":
            from . import _lib
            raise _lib.PamnetError("DeviceDataset keeps the dataset in GPU memory: pass a CUDA device (no CPU fallback)")
        arr = lambda v, dt: np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v, dtype=dt)
        xs = [arr(m.x, np.float32).reshape(-1) for m in molecules]
        ps = [arr(m.pos, np.float32).reshape(-1, 3) for m in molecules]
        es = [arr(m.edge_index, np.int64).reshape(2, -1) for m in molecules]
        for k, (x, p, e) in enumerate(zip(xs, ps, es)):
            if p.shape[0] != x.shape[0] or (e.size and (e.min() < 0 or e.max() >= x.shape[0])):
                raise ValueError(f"molecule {k}: pos / edge_index do not match its {x.shape[0]} atoms")
        self.n_atoms = np.array([x.shape[0] for x in xs], dtype=np.int64)
        self.n_bonds = np.array([e.shape[1] for e in es], dtype=np.int64)
        self.node_ptr_host = np.concatenate([[0], np.cumsum(self.n_atoms)]).astype(np.int64)
        self.edge_ptr_host = np.concatenate([[0], np.cumsum(self.n_bonds)]).astype(np.int64)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.device = device
        self.x_all = to(np.concatenate(xs) if xs else np.zeros(0, np.float32))
        self.pos_all = to(np.concatenate(ps) if ps else np.zeros((0, 3), np.float32))
        self.ei_all = to(np.concatenate(es, axis=1) if es else np.zeros((2, 0), np.int64))
        self.y_all = to(np.array([float(m.y) for m in molecules], dtype=np.float32))
        self.node_ptr, self.edge_ptr = to(self.node_ptr_host), to(self.edge_ptr_host)

    def __len__(self):
        return int(self.n_atoms.shape[0])

    def table(self, ids):
        """Host side of a batch: the [3, G] table and the batch's atom / bond totals."""
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size == 0 or ids.min() < 0 or ids.max() >= len(self):
            raise IndexError("molecule ids out of range")
        na, nb = self.n_atoms[ids], self.n_bonds[ids]
        n0 = np.concatenate([[0], np.cumsum(na)[:-1]])
        e0 = np.concatenate([[0], np.cumsum(nb)[:-1]])
        return np.stack([ids, n0, e0]).astype(np.int64), int(na.sum()), int(nb.sum())

    def batch(self, ids):
        from . import _lib
        lib = _lib.load()
        tab, n, e = self.table(ids)
        g, dev = tab.shape[1], self.device
        tab_dev = torch.from_numpy(tab).pin_memory().to(dev, non_blocking=True)
        x = torch.empty(n, dtype=torch.float32, device=dev)
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        ei = torch.empty((2, e), dtype=torch.int64, device=dev)
        bvec = torch.empty(n, dtype=torch.int64, device=dev)
        y = torch.empty(g, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):            # the library launches on the current device
            _lib.check(lib.pamnet_collate(tab_dev.data_ptr(), g, self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(),
                                          self.x_all.data_ptr(), self.pos_all.data_ptr(), self.ei_all.data_ptr(),
                                          int(self.ei_all.shape[1]), self.y_all.data_ptr(), e, x.data_ptr(), pos.data_ptr(),
                                          ei.data_ptr(), bvec.data_ptr(), y.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "collate")
        return Batch(x=x, pos=pos, edge_index=ei, batch=bvec, y=y)
