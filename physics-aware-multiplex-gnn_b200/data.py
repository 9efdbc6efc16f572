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
