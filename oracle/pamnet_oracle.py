"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch, fp32 or fp64) of the PAMNet
forward pass, written functionally over a reference-keyed ``state_dict``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
leg may import this file; the product path (the CUDA extension behind
``pamnet_b200``) never does, and fails loudly when its shared library is missing.

Pinned against: the reference's own ``models.py`` imported verbatim (through
``oracle/ref_shim.py``) in the build container -- ``tests/test_oracle_vs_reference.py`` --
and the golden vectors under ``tests/golden`` that the same verbatim import produced
(``tests/golden/make_golden.py``).  The third-party graph ops underneath are
parity-unpinned by the reference itself (see ``oracle/graph_ops.py``).

Gradients come from torch autograd over these functions (the reference has no
hand-written backward either: main_qm9.py:110).
"""
import math

import numpy as np
import torch

from . import graph_ops as G

NUM_SPHERICAL, NUM_RADIAL, NUM_RBF, ENVELOPE_P = 7, 6, 16, 5


# ----------------------------------------------------------------------------
# basis constants (utils/sbf.py:13-26 zeros, :41-49 normalisers, :94-139 Y_l0)
# ----------------------------------------------------------------------------
def sbf_constants(n=NUM_SPHERICAL, k=NUM_RADIAL):
    """zeros z[l][m] of j_l (stored as float32 like utils/sbf.py:15) and
    N[l][m] = (0.5*j_{l+1}(z)^2)^-1/2 evaluated in double on the rounded zeros."""
    from scipy.optimize import brentq
    from scipy.special import spherical_jn

    zeros = np.zeros((n, k), dtype=np.float32)
    zeros[0] = np.arange(1, k + 1) * np.pi
    brackets = np.arange(1, k + n) * np.pi
    for l in range(1, n):
        found = [brentq(lambda r: spherical_jn(l, r), float(brackets[m]), float(brackets[m + 1]))
                 for m in range(k + n - 1 - l)]
        # the reference keeps the running brackets in a float32 array (sbf.py:17,23-24)
        brackets = np.asarray(found, dtype=np.float32)
        zeros[l] = brackets[:k]
    z64 = zeros.astype(np.float64)
    norm = np.zeros((n, k))
    for l in range(n):
        norm[l] = 1.0 / np.sqrt(0.5 * spherical_jn(l + 1, z64[l]) ** 2)
    return zeros, norm


def zonal_harmonic_coeffs(n=NUM_SPHERICAL):
    """Y_l0(theta) = sum_p c[l][p] cos(theta)^p  (utils/sbf.py:62-66,69-79,129-131)."""
    leg = [np.array([1.0]), np.array([0.0, 1.0])]
    for j in range(2, n):
        a = np.zeros(j + 1)
        a[1:] += (2 * j - 1) * leg[j - 1]
        a[:j - 1] -= (j - 1) * leg[j - 2]
        leg.append(a / j)
    out = np.zeros((n, n))
    for l in range(n):
        out[l, :l + 1] = math.sqrt((2 * l + 1) / (4 * math.pi)) * leg[l]
    return out


def spherical_bessel_upto(lmax, x):
    """j_0..j_lmax by upward recurrence from the closed forms of j_0, j_1 (utils/sbf.py:29-38
    generates the equivalent closed forms symbolically)."""
    s, c = torch.sin(x), torch.cos(x)
    js = [s / x, (s / x - c) / x]
    for l in range(1, lmax):
        js.append((2 * l + 1) / x * js[l] - js[l - 1])
    return js[:lmax + 1]


def envelope(x, p=ENVELOPE_P):
    """layers/basic.py:36-51 (exponent p, not p-1)."""
    a, b, c = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
    xp = x.pow(p)
    val = 1.0 / x + a * xp + b * xp * x + c * xp * x * x
    return torch.where(x < 1, val, torch.zeros_like(x))


def bessel_rbf(dist, freq, cutoff):
    """layers/basic.py:74-76."""
    x = dist.unsqueeze(-1) / cutoff
    return envelope(x) * torch.sin(freq * x)


def spherical_basis(dist, angle, gather, cutoff, consts=None):
    """layers/basic.py:107-116: [T, 42] with column l*6+m."""
    zeros, norm = consts if consts is not None else sbf_constants()
    ycoef = zonal_harmonic_coeffs()
    x = dist / cutoff
    z = torch.as_tensor(zeros.astype(np.float64), dtype=dist.dtype)
    nrm = torch.as_tensor(norm, dtype=dist.dtype)
    radial = []
    for l in range(NUM_SPHERICAL):
        arg = z[l][None, :] * x[:, None]                         # [E, 6]
        radial.append(nrm[l][None, :] * spherical_bessel_upto(l, arg)[l])
    radial = torch.stack(radial, 1) * envelope(x)[:, None, None]  # [E, 7, 6]
    ct = torch.cos(angle)
    yc = torch.as_tensor(ycoef, dtype=dist.dtype)
    powers = torch.stack([ct ** p for p in range(NUM_SPHERICAL)], 1)   # [T, 7]
    cbf = powers @ yc.T                                                  # [T, 7]
    return (radial[gather] * cbf[:, :, None]).reshape(-1, NUM_SPHERICAL * NUM_RADIAL)


# ----------------------------------------------------------------------------
# dense blocks (layers/basic.py:11-33)
# ----------------------------------------------------------------------------
def silu(x):
    return x * torch.sigmoid(x)


def _lin(sd, key, x, act=True):
    y = x @ sd[key + ".weight"].T
    if key + ".bias" in sd:
        y = y + sd[key + ".bias"]
    return silu(y) if act else y


def _mlp(sd, prefix, x, n):
    """MLP([..]) of n Linear+SiLU stages: keys prefix.{s}.0.{weight,bias} (basic.py:19-22)."""
    for s in range(n):
        x = _lin(sd, f"{prefix}.{s}.0", x)
    return x


def _res(sd, prefix, x):
    return x + _mlp(sd, prefix + ".mlp", x, 2)


def _tail(sd, p, h, res_x):
    """Update block + readout heads shared by both layer kinds
    (global_message_passing.py:39-48, local_message_passing.py:55-64)."""
    x = _mlp(sd, p + ".mlp_x2", h, 1)
    x = _res(sd, p + ".res1", x) + res_x
    x = _res(sd, p + ".res2", x)
    x = _res(sd, p + ".res3", x)
    o = _mlp(sd, p + ".mlp_out", x, 3)
    att = o @ sd[p + ".W"]
    out = o @ sd[p + ".W_out.weight"].T + sd[p + ".W_out.bias"]
    return x, out, att


def global_layer(sd, p, x, edge_attr, edge_index, flow="source_to_target"):
    """Global_MessagePassing.forward/message (global_message_passing.py:33-56) with PyG's
    propagate: x_i = x[edge_index[i]], aggregation at edge_index[i], (i, j) = (1, 0) for
    source_to_target and (0, 1) for target_to_source."""
    i, j = (0, 1) if flow == "target_to_source" else (1, 0)
    x1 = _mlp(sd, p + ".mlp_x1", x, 1)
    m = torch.cat((x1[edge_index[i]], x1[edge_index[j]], edge_attr), -1)
    m = _mlp(sd, p + ".mlp_m", m, 1) * (edge_attr @ sd[p + ".W_edge_attr.weight"].T)
    h = x1 + G.segment_sum(m, edge_index[i], x.shape[0])
    return _tail(sd, p, h, x)


def local_layer(sd, p, x, rbf, sbf2, sbf1, idx_kj, idx_ji, idx_jj_pair, idx_ji_pair, edge_index,
                two_hop=True):
    """Local_MessagePassing.forward (local_message_passing.py:36-66); two_hop=False is
    Local_MessagePassing_s (:98-123, weights under mlp_m_jj)."""
    j, i = edge_index
    x1 = _mlp(sd, p + ".mlp_x1", x, 1)
    m = torch.cat((x1[i], x1[j], rbf), -1)
    m_ji = _mlp(sd, p + ".mlp_m_ji", m, 1)
    nb_key = ".mlp_m_kj" if two_hop else ".mlp_m_jj"
    m_nb = _mlp(sd, p + nb_key, m, 1) * (rbf @ sd[p + ".lin_rbf.weight"].T)
    if two_hop:
        gather = torch.cat((idx_kj, idx_jj_pair))
        scatter = torch.cat((idx_ji, idx_ji_pair))
        sbf = torch.cat((sbf2, sbf1))
    else:
        gather, scatter, sbf = idx_jj_pair, idx_ji_pair, sbf1
    m_other = G.segment_sum(m_nb[gather] * _mlp(sd, p + ".mlp_sbf", sbf, 2), scatter, m.shape[0])
    m = (rbf @ sd[p + ".lin_rbf_out.weight"].T) * (m_ji + m_other)
    h = x1 + G.segment_sum(m, i, x.shape[0])
    return _tail(sd, p, h, x)


# ----------------------------------------------------------------------------
# whole model (models.py:100-224, :285-353)
# ----------------------------------------------------------------------------
class Graph:
    """Everything layer-invariant that PAMNet.forward derives from the batch."""
    pass


def build_graph(dataset, data, cutoff_l, cutoff_g, simple=False):
    """models.py:104-177 (PAMNet) / :289-320 (PAMNet_s): edges, index vectors, geometry."""
    g = Graph()
    kind = "rna" if dataset[:3].lower() == "rna" else dataset
    batch = data.batch
    if kind == "QM9":
        pos = data.pos
        row, col = G.radius_pairs(pos, pos, cutoff_g, batch, batch, 500 if simple else 1000)
        g.edge_index_g = G.drop_self_loops(torch.stack([row, col]))
        g.edge_index_l = G.drop_self_loops(data.edge_index)
    elif kind == "PDBbind":
        xr = data.x.unsqueeze(-1) if data.x.dim() == 1 else data.x
        pos = xr[:, :3].contiguous()
        row, col = G.radius_pairs(pos, pos, cutoff_g, batch, batch, 1000)
        g.edge_index_g = G.drop_self_loops(torch.stack([row, col]))
        keep = G.edge_lengths(g.edge_index_g, pos) <= cutoff_l
        g.edge_index_l = G.drop_self_loops(g.edge_index_g[:, keep])
    elif kind == "rna":
        xr = data.x.unsqueeze(-1) if data.x.dim() == 1 else data.x
        pos = xr[:, :3].contiguous()
        row, col = G.knn_pairs(pos, pos, 50, batch, batch)
        knn = G.drop_self_loops(torch.stack([row, col]))
        d = G.edge_lengths(knn, pos)
        g.edge_index_g = knn[:, d <= cutoff_g]
        g.edge_index_l = knn[:, d <= cutoff_l]
    else:
        raise ValueError("Invalid dataset.")
    g.pos = pos
    g.dist_g = G.edge_lengths(g.edge_index_g, pos)
    g.dist_l = G.edge_lengths(g.edge_index_l, pos)
    (g.idx_i, g.idx_j, g.idx_k, g.idx_kj, g.idx_ji, g.idx_i_pair, g.idx_j1_pair, g.idx_j2_pair,
     g.idx_jj_pair, g.idx_ji_pair) = G.triplet_indices(g.edge_index_l, pos.shape[0])
    g.angle2 = G.bond_angle(pos, g.idx_i, g.idx_j, g.idx_k)
    g.angle1 = G.bond_angle(pos, g.idx_i_pair, g.idx_j1_pair, g.idx_j2_pair)
    return g


def forward(sd, cfg, data, simple=False, consts=None, return_parts=False):
    """PAMNet.forward (models.py:100-224); simple=True is PAMNet_s.forward (:285-353).

    ``sd``: dict keyed like the reference state_dict (tensors may require grad);
    ``cfg``: object with dataset, dim, n_layer, cutoff_l, cutoff_g, flow.
    """
    kind = "rna" if cfg.dataset[:3].lower() == "rna" else cfg.dataset
    if simple and kind != "QM9":
        raise ValueError("Invalid dataset. The current PAMNet_s is only for QM9 experiments.")
    g = build_graph(cfg.dataset, data, cfg.cutoff_l, cfg.cutoff_g, simple)
    dtype = sd["embeddings"].dtype
    pos = g.pos.to(dtype)
    if pos.dtype != g.pos.dtype:       # fp64 rung: redo the geometry in double on the same graph
        g.dist_g, g.dist_l = G.edge_lengths(g.edge_index_g, pos), G.edge_lengths(g.edge_index_l, pos)
        g.angle2 = G.bond_angle(pos, g.idx_i, g.idx_j, g.idx_k)
        g.angle1 = G.bond_angle(pos, g.idx_i_pair, g.idx_j1_pair, g.idx_j2_pair)
    xr = data.x
    if kind == "QM9":
        x = sd["embeddings"][xr.long()]
    elif kind == "PDBbind":
        xr = xr.unsqueeze(-1) if xr.dim() == 1 else xr
        x = xr[:, 3:].to(dtype) @ sd["init_linear.weight"].T
        sign = torch.where(xr[:, 0] > 40.0, -1.0, 1.0).to(dtype)
    else:
        xr = xr.unsqueeze(-1) if xr.dim() == 1 else xr
        x = sd["embeddings"][xr[:, -1].long()]

    consts = consts if consts is not None else sbf_constants()
    rbf_l = bessel_rbf(g.dist_l, sd["rbf_l.freq"], cfg.cutoff_l)
    rbf_g = bessel_rbf(g.dist_g, sd["rbf_g.freq"], cfg.cutoff_g)
    sbf1 = spherical_basis(g.dist_l, g.angle1, g.idx_jj_pair, cfg.cutoff_l, consts)
    e_l = _mlp(sd, "mlp_rbf_l", rbf_l, 1)
    e_g = _mlp(sd, "mlp_rbf_g", rbf_g, 1)
    if simple:
        s1, s2 = _mlp(sd, "mlp_sbf", sbf1, 1), None
    else:
        sbf2 = spherical_basis(g.dist_l, g.angle2, g.idx_kj, cfg.cutoff_l, consts)
        s1, s2 = _mlp(sd, "mlp_sbf1", sbf1, 1), _mlp(sd, "mlp_sbf2", sbf2, 1)

    outs_g, outs_l, atts_g, atts_l, x_halves = [], [], [], [], []
    for l in range(cfg.n_layer):
        x, o, a = global_layer(sd, f"global_layer.{l}", x, e_g, g.edge_index_g,
                               getattr(cfg, "flow", "source_to_target"))
        outs_g.append(o), atts_g.append(a), x_halves.append(x)
        x, o, a = local_layer(sd, f"local_layer.{l}", x, e_l, s2, s1, g.idx_kj, g.idx_ji,
                              g.idx_jj_pair, g.idx_ji_pair, g.edge_index_l, two_hop=not simple)
        outs_l.append(o), atts_l.append(a), x_halves.append(x)

    # fusion (models.py:206-213): softmax over the {global, local} pair of every layer
    att = torch.stack((torch.stack(atts_g), torch.stack(atts_l)), -1).squeeze(2)   # [L, N, 2]
    out = torch.stack((torch.stack(outs_g), torch.stack(outs_l)), -1).squeeze(2)
    w = torch.softmax(torch.nn.functional.leaky_relu(att, 0.2), -1)
    node = (out * w).sum(-1).sum(0)                                                  # [N]
    n_graphs = int(data.batch.max()) + 1
    if kind == "PDBbind":
        node = node * sign
    pooled = G.segment_sum(node, data.batch, n_graphs)
    if kind == "rna":
        cnt = torch.bincount(data.batch, minlength=n_graphs).clamp(min=1).to(dtype)
        pooled = pooled / cnt
    if return_parts:
        return pooled, dict(graph=g, rbf_l=rbf_l, rbf_g=rbf_g, sbf1=sbf1, e_g=e_g, e_l=e_l, s1=s1, s2=s2,
                            x_last=x, x_halves=x_halves, att=att, out=out)
    return pooled


# ----------------------------------------------------------------------------
# parameter table (key order == reference state_dict order)
# ----------------------------------------------------------------------------
def param_shapes(cfg, simple=False):
    D = cfg.dim
    rna = cfg.dataset[:3].lower() == "rna"
    t = [("embeddings", (3 if rna else 5, D))]
    if not rna and not simple:
        t.append(("init_linear.weight", (D, 18)))
    t += [("rbf_g.freq", (NUM_RBF,)), ("rbf_l.freq", (NUM_RBF,))]

    def lin(name, o, i, bias=True):
        return [(name + ".weight", (o, i))] + ([(name + ".bias", (o,))] if bias else [])

    nsbf = NUM_SPHERICAL * NUM_RADIAL
    t += lin("mlp_rbf_g.0.0", D, NUM_RBF) + lin("mlp_rbf_l.0.0", D, NUM_RBF)
    if simple:
        t += lin("mlp_sbf.0.0", D, nsbf)
    else:
        t += lin("mlp_sbf1.0.0", D, nsbf) + lin("mlp_sbf2.0.0", D, nsbf)

    def res(p):
        return sum((lin(f"{p}.res{r}.mlp.{s}.0", D, D) for r in (1, 2, 3) for s in (0, 1)), [])

    def heads(p):
        return sum((lin(f"{p}.mlp_out.{s}.0", D, D) for s in range(3)), []) + lin(p + ".W_out", 1, D)

    for l in range(cfg.n_layer):
        p = f"global_layer.{l}"
        t += [(p + ".W", (D, 1))] + lin(p + ".mlp_x1.0.0", D, D) + lin(p + ".mlp_x2.0.0", D, D) + res(p)
        t += lin(p + ".mlp_m.0.0", D, 3 * D) + lin(p + ".W_edge_attr", D, D, False) + heads(p)
    nb = "mlp_m_jj" if simple else "mlp_m_kj"
    for l in range(cfg.n_layer):
        p = f"local_layer.{l}"
        t += [(p + ".W", (D, 1))] + lin(p + ".mlp_x1.0.0", D, D) + lin(p + ".mlp_m_ji.0.0", D, 3 * D)
        t += lin(f"{p}.{nb}.0.0", D, 3 * D) + lin(p + ".mlp_sbf.0.0", D, D) + lin(p + ".mlp_sbf.1.0", D, D)
        t += lin(p + ".lin_rbf", D, D, False) + res(p) + lin(p + ".lin_rbf_out", D, D, False)
        t += lin(p + ".mlp_x2.0.0", D, D) + heads(p)
    return t


def init_state_dict(cfg, seed=0, simple=False, dtype=torch.float32):
    """Reference-style initial values (not its RNG stream): nn.Linear default
    U(+-1/sqrt(fan_in)) (weight: kaiming a=sqrt(5) == same bound), glorot on W
    (global_message_passing.py:31), embeddings U(+-sqrt(3)) (models.py:58-60), freq = n*pi
    (basic.py:69-72)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg, simple):
        if name == "embeddings":
            bound = math.sqrt(3)
        elif name.endswith(".freq"):
            sd[name] = (torch.arange(1, NUM_RBF + 1, dtype=torch.float64) * math.pi).to(dtype)
            continue
        elif name.endswith(".W"):
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        elif name.endswith(".bias"):
            bound = 1.0 / math.sqrt(sd[name[:-5] + ".weight"].shape[1])
        else:
            bound = 1.0 / math.sqrt(shape[1])
        sd[name] = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return sd
