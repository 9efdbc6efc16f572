"""Module tree of the reference layers (parameter containers with the reference's names, shapes and
initialisation) -- layers/basic.py, layers/global_message_passing.py, layers/local_message_passing.py.

The arithmetic of a whole model step runs in the fused CUDA path (models.py -> libpamnet_sm100.so); the
``forward`` methods here are the stand-alone layer surface: same signatures as the reference
layers, composed from the operator kernels in ops.py (ops.linear / ops.scatter are autograd functions over the C ABI's
GEMM and segment-sum kernels), differentiable w.r.t. every input and parameter (tests/test_gpu_layers.py).
"""
import math

import torch
import torch.nn as nn

from . import ops


class SiLU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def MLP(channels):
    """Sequential of Sequential(Linear, SiLU): keys '<i>.0.weight' / '<i>.0.bias' (layers/basic.py:19-22)."""
    return nn.Sequential(*[nn.Sequential(nn.Linear(channels[i - 1], channels[i]), SiLU())
                           for i in range(1, len(channels))])


def _run_mlp(mlp, x):
    for stage in mlp:
        x = ops.linear(x, stage[0].weight, stage[0].bias, silu=True)
    return x


class Res(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.mlp = MLP([dim, dim, dim])

    def forward(self, x):
        return _run_mlp(self.mlp, x) + x


class BesselBasisLayer(nn.Module):
    """layers/basic.py:59-76; freq is learnable, initialised to n*pi."""

    def __init__(self, num_radial, cutoff, envelope_exponent=5):
        super().__init__()
        if num_radial != 16 or envelope_exponent != 5:
            raise ValueError("the CUDA path implements the reference configuration: 16 radial functions, exponent 5")
        self.cutoff = cutoff
        self.freq = nn.Parameter(torch.arange(1, num_radial + 1, dtype=torch.float32) * math.pi)

    def forward(self, dist):
        return ops.bessel_rbf(dist, self.freq, self.cutoff)


class SphericalBasisLayer(nn.Module):
    """layers/basic.py:79-116; no parameters."""

    def __init__(self, num_spherical, num_radial, cutoff=5.0, envelope_exponent=5):
        super().__init__()
        if (num_spherical, num_radial, envelope_exponent) != (7, 6, 5):
            raise ValueError("the CUDA path implements the reference configuration: 7 x 6 basis, exponent 5")
        self.num_spherical, self.num_radial, self.cutoff = num_spherical, num_radial, cutoff

    def forward(self, dist, angle, idx_kj):
        return ops.spherical_basis(dist, angle, idx_kj, self.cutoff)


def _glorot(t):
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


def _update_and_heads(layer, h, res_x):
    x = _run_mlp(layer.mlp_x2, h)
    x = layer.res1(x) + res_x
    x = layer.res2(x)
    x = layer.res3(x)
    o = _run_mlp(layer.mlp_out, x)
    att = ops.linear(o, layer.W.t().contiguous()).unsqueeze(0)
    out = ops.linear(o, layer.W_out.weight, layer.W_out.bias).unsqueeze(0)
    return x, out, att


class Global_MessagePassing(nn.Module):
    """global_message_passing.py:9-60."""

    def __init__(self, config):
        super().__init__()
        self.dim = d = config.dim
        self.flow = getattr(config, "flow", "source_to_target")
        self.mlp_x1 = MLP([d, d])
        self.mlp_x2 = MLP([d, d])
        self.res1, self.res2, self.res3 = Res(d), Res(d), Res(d)
        self.mlp_m = MLP([3 * d, d])
        self.W_edge_attr = nn.Linear(d, d, bias=False)
        self.mlp_out = MLP([d, d, d, d])
        self.W_out = nn.Linear(d, 1)
        self.W = nn.Parameter(torch.empty(d, 1))
        _glorot(self.W)

    def forward(self, x, edge_attr, edge_index):
        i, j = (0, 1) if self.flow == "target_to_source" else (1, 0)
        x1 = _run_mlp(self.mlp_x1, x)
        m = torch.cat((x1[edge_index[i]], x1[edge_index[j]], edge_attr), -1)
        m = _run_mlp(self.mlp_m, m) * ops.linear(edge_attr, self.W_edge_attr.weight)
        h = x1 + ops.scatter(m, edge_index[i], dim_size=x.shape[0])
        return _update_and_heads(self, h, x)


class _LocalBase(nn.Module):
    def _init(self, config, nb_name):
        self.dim = d = config.dim
        self.mlp_x1 = MLP([d, d])
        self.mlp_m_ji = MLP([3 * d, d])
        setattr(self, nb_name, MLP([3 * d, d]))
        self.mlp_sbf = MLP([d, d, d])
        self.lin_rbf = nn.Linear(d, d, bias=False)
        self.res1, self.res2, self.res3 = Res(d), Res(d), Res(d)
        self.lin_rbf_out = nn.Linear(d, d, bias=False)
        self.mlp_x2 = MLP([d, d])
        self.mlp_out = MLP([d, d, d, d])
        self.W_out = nn.Linear(d, 1)
        self.W = nn.Parameter(torch.empty(d, 1))
        _glorot(self.W)

    def _forward(self, nb_mlp, x, rbf, sbf, idx, idx_scatter, edge_index):
        j, i = edge_index
        x1 = _run_mlp(self.mlp_x1, x)
        m = torch.cat([x1[i], x1[j], rbf], -1)
        m_ji = _run_mlp(self.mlp_m_ji, m)
        m_nb = _run_mlp(nb_mlp, m) * ops.linear(rbf, self.lin_rbf.weight)
        m_other = ops.scatter(m_nb[idx] * _run_mlp(self.mlp_sbf, sbf), idx_scatter, dim_size=m.shape[0])
        m = ops.linear(rbf, self.lin_rbf_out.weight) * (m_ji + m_other)
        h = x1 + ops.scatter(m, i, dim_size=x.shape[0])
        return _update_and_heads(self, h, x)


class Local_MessagePassing(_LocalBase):
    """local_message_passing.py:9-66."""

    def __init__(self, config):
        super().__init__()
        self._init(config, "mlp_m_kj")

    def forward(self, x, rbf, sbf2, sbf1, idx_kj, idx_ji, idx_jj_pair, idx_ji_pair, edge_index):
        return self._forward(self.mlp_m_kj, x, rbf, torch.cat((sbf2, sbf1)), torch.cat((idx_kj, idx_jj_pair)),
                             torch.cat((idx_ji, idx_ji_pair)), edge_index)


class Local_MessagePassing_s(_LocalBase):
    """local_message_passing.py:69-123."""

    def __init__(self, config):
        super().__init__()
        self._init(config, "mlp_m_jj")

    def forward(self, x, rbf, sbf, idx_jj_pair, idx_ji_pair, edge_index):
        return self._forward(self.mlp_m_jj, x, rbf, sbf, idx_jj_pair, idx_ji_pair, edge_index)
