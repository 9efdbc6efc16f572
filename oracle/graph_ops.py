"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the graph-structural steps of the
PAMNet hot path.  Never imported by the product path (only tests/, smoke() and
bench.py's cpu_baseline/reference leg).

Each function cites the reference line it follows.  The reference delegates most of
these to third-party packages that are absent here (torch_cluster 1.5.4,
torch_sparse 0.6.0, torch_scatter 2.0.4, torch-geometric 1.4.2;
/root/reference/requirements.txt:8-11) and ships no tests, so **parity at these
boundaries is unpinned**; the canonical semantics are stated here once and the CUDA
kernels are graded against them:

* ``radius``: pairs (query q, neighbour n) of the same example with
  d2 = ((dx*dx)+(dy*dy))+(dz*dz) evaluated in fp32 *without FMA contraction* and
  ``d2 <= fl32(r)*fl32(r)``; q itself included; per query the first
  ``max_num_neighbors`` in ascending n; output ordered by (q, n).
* ``knn``: per query the k smallest d2 (same formula), ties to the lower index, ascending d2.
"""
import torch


def _segments(batch):
    """[start, end) of every example in a non-decreasing batch vector."""
    if batch.numel() == 0:
        return []
    assert bool((batch[1:] >= batch[:-1]).all()), "batch vector must be non-decreasing"
    n_graphs = int(batch[-1]) + 1
    counts = torch.bincount(batch, minlength=n_graphs)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    return [(int(s), int(e)) for s, e in zip(starts, ends) if e > s]


def _d2_block(q, p):
    """Canonical squared distance, fp32 op order ((dx*dx)+(dy*dy))+(dz*dz)."""
    dx = q[:, None, 0] - p[None, :, 0]
    dy = q[:, None, 1] - p[None, :, 1]
    dz = q[:, None, 2] - p[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def radius_pairs(x, y, r, batch_x, batch_y, max_num_neighbors=32):
    """torch_cluster.radius as called at models.py:110,128,301 (x is y there)."""
    assert x is y or (x.shape == y.shape and torch.equal(batch_x, batch_y)), \
        "oracle restates the self-query form used by the reference"
    r2 = (torch.tensor(float(r), dtype=x.dtype) * torch.tensor(float(r), dtype=x.dtype))
    rows, cols = [], []
    for s, e in _segments(batch_x):
        d2 = _d2_block(y[s:e], x[s:e])
        hit = d2 <= r2
        if max_num_neighbors < e - s:
            rank = torch.cumsum(hit.to(torch.long), 1)
            hit = hit & (rank <= max_num_neighbors)
        rq, cn = hit.nonzero(as_tuple=True)
        rows.append(rq + s)
        cols.append(cn + s)
    if not rows:
        z = torch.zeros(0, dtype=torch.long)
        return z, z.clone()
    return torch.cat(rows), torch.cat(cols)


def knn_pairs(x, y, k, batch_x, batch_y):
    """torch_cluster.knn as called at models.py:143 (x is y there)."""
    rows, cols = [], []
    for s, e in _segments(batch_x):
        n = e - s
        d2 = _d2_block(y[s:e], x[s:e])
        order = torch.sort(d2, dim=1, stable=True).indices[:, :min(k, n)]
        rows.append((torch.arange(n)[:, None] + s).expand_as(order).reshape(-1))
        cols.append(order.reshape(-1) + s)
    if not rows:
        z = torch.zeros(0, dtype=torch.long)
        return z, z.clone()
    return torch.cat(rows), torch.cat(cols)


def drop_self_loops(edge_index):
    """torch_geometric.utils.remove_self_loops (models.py:63): order-preserving mask."""
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep]


def edge_lengths(edge_index, pos):
    """models.py:64-65: j, i = edge_index; |pos[i]-pos[j]|."""
    j, i = edge_index
    d = pos[i] - pos[j]
    return (d * d).sum(-1).sqrt()


def segment_sum(src, index, n_rows):
    """torch_scatter.scatter(..., reduce='add') (local_message_passing.py:50,54)."""
    out = torch.zeros((n_rows,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def incoming_csr(edge_index, num_nodes):
    """CSR keyed by TARGET node (models.py:72: SparseTensor(row=col, col=row, value=arange)).

    Returns (ptr[num_nodes+1], src[E], eid[E]); within a target the entries are ordered by
    source node id, then by edge id (stable sort of col*num_nodes+row).
    """
    row, col = edge_index
    order = torch.argsort(col * num_nodes + row, stable=True)
    counts = torch.bincount(col, minlength=num_nodes)
    ptr = torch.zeros(num_nodes + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr, row[order], order


def _expand(ptr, keys):
    """For each position p, list CSR row keys[p]: returns (owner position, csr slot)."""
    start = ptr[keys]
    cnt = ptr[keys + 1] - start
    owner = torch.repeat_interleave(torch.arange(keys.numel()), cnt)
    first = torch.cumsum(cnt, 0) - cnt
    slot = start[owner] + (torch.arange(int(cnt.sum())) - first[owner])
    return owner, slot


def triplet_indices(edge_index, num_nodes):
    """PAMNet.indices (models.py:68-98) as two walks over the incoming-edge CSR.

    Two-hop (models.py:74-84): for edge e = (j -> i), every edge (k -> j) with k != i.
    One-hop (models.py:85-96): for edge e = (j -> i), every edge (j' -> i); the reference's
    mask compares the *target* i with j' (models.py:92), which only drops self loops, so the
    pair (e, e) is kept.
    Returns the ten vectors in the reference's order.
    """
    row, col = edge_index
    ptr, src, eid = incoming_csr(edge_index, num_nodes)

    own, slot = _expand(ptr, row)             # edges entering the SOURCE j of e
    idx_i, idx_j, idx_k = col[own], row[own], src[slot]
    keep = idx_i != idx_k
    idx_i, idx_j, idx_k = idx_i[keep], idx_j[keep], idx_k[keep]
    idx_kj, idx_ji = eid[slot][keep], own[keep]

    own, slot = _expand(ptr, col)             # edges entering the TARGET i of e
    idx_i_pair, idx_j1_pair, idx_j2_pair = row[own], col[own], src[slot]
    keep = idx_j1_pair != idx_j2_pair
    idx_i_pair, idx_j1_pair, idx_j2_pair = idx_i_pair[keep], idx_j1_pair[keep], idx_j2_pair[keep]
    idx_ji_pair, idx_jj_pair = own[keep], eid[slot][keep]

    return (idx_i, idx_j, idx_k, idx_kj, idx_ji,
            idx_i_pair, idx_j1_pair, idx_j2_pair, idx_jj_pair, idx_ji_pair)


def bond_angle(pos, a, b, c):
    """models.py:165-177: angle between (pos[b]-pos[a]) and (pos[c]-pos[b]) via atan2(|cross|, dot)."""
    u, v = pos[b] - pos[a], pos[c] - pos[b]
    dot = (u * v).sum(-1)
    cross = torch.linalg.cross(u, v, dim=-1).norm(dim=-1)
    return torch.atan2(cross, dot)
