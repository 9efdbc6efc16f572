"""numpy restatement of the execution plan (csrc/graph.cuh:Plan) from its definition -- test infrastructure only.

The plan is the destination-sorted form of the two edge lists plus the merged triplet lists:

* in-CSR of a graph: slots keyed by the aggregation target, inside a target ordered by (source id, API edge id) --
  SparseTensor's order at models.py:72;
* out-CSR: keyed by the other end, ordered by API edge id; ``opos`` maps its entries to in-CSR slots;
* triplets of local slot k = (j -> i): first the slots into j whose source is not i (two-hop, models.py:74-85), then the
  slots into i whose source is not i (one-hop, models.py:87-96) -- the reference's cat() order
  (layers/local_message_passing.py:38-40); ``tt``: triplet ids grouped by the slot they gather, ascending.
"""
import numpy as np


def _csr(dst, src, n_nodes):
    e = np.arange(dst.shape[0])
    order = np.lexsort((e, src, dst))                      # by (dst, src, eid)
    ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n_nodes), out=ptr[1:])
    pos_of = np.empty_like(order)
    pos_of[order] = np.arange(order.shape[0])
    oorder = np.lexsort((e, src))                          # out-CSR: by (src, eid)
    optr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n_nodes), out=optr[1:])
    return dict(ptr=ptr, eid=order, src=src[order], dst=dst[order], optr=optr, opos=pos_of[oorder])


def _edge_len(pos, a, b):
    d = pos[a] - pos[b]
    return np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)


def _angle(pos, a, b, c):
    u = (pos[b] - pos[a]).astype(np.float64)
    v = (pos[c] - pos[b]).astype(np.float64)
    return np.arctan2(np.linalg.norm(np.cross(u, v), axis=-1), (u * v).sum(-1))


def build_plan(pos, batch, n_graphs, edge_index_g, edge_index_l, g_dst_row, two_hop=True):
    """All plan arrays as a dict of numpy arrays (int64 / float32 / float64 angles)."""
    pos = np.asarray(pos, dtype=np.float32)
    batch = np.asarray(batch, dtype=np.int64)
    eg = np.asarray(edge_index_g, dtype=np.int64)
    el = np.asarray(edge_index_l, dtype=np.int64)
    n = pos.shape[0]
    g = _csr(eg[g_dst_row], eg[1 - g_dst_row], n)
    l = _csr(el[1], el[0], n)                              # local: i = edge_index[1] (local_message_passing.py:37)
    out = {"n2g": batch.copy(), "gptr": np.searchsorted(batch, np.arange(n_graphs + 1), side="left")}
    for k, v in g.items():
        out["g_" + k] = v
    for k, v in l.items():
        out["l_" + k] = v
    out["dist_g"] = _edge_len(pos, g["dst"], g["src"])
    out["dist_l"] = _edge_len(pos, l["dst"], l["src"])
    n_l = el.shape[1]
    t_split, t_cnt, t_gather, t_owner, tri = [], [], [], [], []
    lp, ls, ld = l["ptr"], l["src"], l["dst"]
    for k in range(n_l):
        j, i = int(ls[k]), int(ld[k])
        two = [p for p in range(lp[j], lp[j + 1]) if ls[p] != i] if two_hop else []
        one = [p for p in range(lp[i], lp[i + 1]) if ls[p] != i]
        t_split.append(len(two))
        t_cnt.append(len(two) + len(one))
        for p in two:
            t_gather.append(p); t_owner.append(k); tri.append((i, j, int(ls[p])))
        for p in one:
            t_gather.append(p); t_owner.append(k); tri.append((j, i, int(ls[p])))
    t_gather = np.asarray(t_gather, dtype=np.int64)
    out["t_split"] = np.asarray(t_split, dtype=np.int64)
    out["t_cnt"] = np.asarray(t_cnt, dtype=np.int64)
    out["t_ptr"] = np.concatenate([[0], np.cumsum(out["t_cnt"])]).astype(np.int64)
    out["t_gather"] = t_gather
    out["t_owner"] = np.asarray(t_owner, dtype=np.int64)
    out["tt_t"] = np.argsort(t_gather, kind="stable")
    out["tt_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(t_gather, minlength=n_l))]).astype(np.int64)
    tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
    out["t_angle"] = _angle(pos, tri[:, 0], tri[:, 1], tri[:, 2])
    # coincident atoms make a leg of the angle zero: atan2(0, +-0) is 0 or pi depending on the sign of the zero
    out["t_angle_defined"] = (np.abs(pos[tri[:, 1]] - pos[tri[:, 0]]).sum(-1) > 0) & (np.abs(pos[tri[:, 2]] - pos[tri[:, 1]]).sum(-1) > 0)
    out["n_t2"] = int(out["t_split"].sum())
    out["n_t1"] = int(out["t_cnt"].sum() - out["t_split"].sum())
    return out
