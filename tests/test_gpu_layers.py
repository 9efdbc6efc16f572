"""GPU: the stand-alone layer surface -- Global_MessagePassing / Local_MessagePassing /
Local_MessagePassing_s with the reference's signatures (global_message_passing.py:33, local_message_passing.py:36,98),
forward AND backward (every input and every parameter) against the oracle's layer functions in fp64."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _graph(n_graphs=4, seed=2):
    from pamnet_b200.data import synthetic_qm9_batch
    from oracle import pamnet_oracle as O
    b = synthetic_qm9_batch(n_graphs, seed=seed)
    g = O.build_graph("QM9", b, 5.0, 5.0)
    return b, g


def _sd_of(layer, prefix):
    return {prefix + "." + k: v.detach().double().cpu().requires_grad_(True) for k, v in layer.state_dict(keep_vars=True).items()}


def _check(outs, refs, ins, ref_ins, layer, sd, prefix, tol=2e-5):
    for o, r in zip(outs, refs):
        assert rel_err(o.detach().reshape(-1), r.detach().reshape(-1)) < tol
    w = [torch.randn_like(o) for o in outs]
    sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
    sum((r.reshape(wi.shape) * wi.double().cpu()).sum() for r, wi in zip(refs, w)).backward()
    for a, b in zip(ins, ref_ins):
        assert a.grad is not None and rel_err(a.grad, b.grad) < tol
    for k, p in layer.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad.reshape(-1), sd[prefix + "." + k].grad.reshape(-1)) < tol, k


@pytest.mark.parametrize("flow", ["source_to_target", "target_to_source"])
def test_global_layer_forward_backward(flow):
    from pamnet_b200 import Config
    from pamnet_b200.layers import Global_MessagePassing
    from oracle import pamnet_oracle as O
    torch.manual_seed(0)
    b, g = _graph()
    d = 64
    layer = Global_MessagePassing(Config("QM9", d, 1, 5.0, 5.0, flow)).cuda()
    sd = _sd_of(layer, "L")
    ei = g.edge_index_g
    x = torch.randn(b.x.shape[0], d, device="cuda", requires_grad=True)
    ea = torch.randn(ei.shape[1], d, device="cuda", requires_grad=True)
    x64, ea64 = x.detach().double().cpu().requires_grad_(True), ea.detach().double().cpu().requires_grad_(True)
    outs = layer(x, ea, ei.cuda())
    refs = O.global_layer(sd, "L", x64, ea64, ei, flow)
    _check(outs, refs, [x, ea], [x64, ea64], layer, sd, "L")


@pytest.mark.parametrize("simple", [False, True])
def test_local_layer_forward_backward(simple):
    from pamnet_b200 import Config
    from pamnet_b200.layers import Local_MessagePassing, Local_MessagePassing_s
    from oracle import pamnet_oracle as O
    torch.manual_seed(1)
    b, g = _graph()
    d = 64
    layer = (Local_MessagePassing_s if simple else Local_MessagePassing)(Config("QM9", d, 1, 5.0, 5.0)).cuda()
    sd = _sd_of(layer, "L")
    ei = g.edge_index_l
    n_e = ei.shape[1]
    x = torch.randn(b.x.shape[0], d, device="cuda", requires_grad=True)
    rbf = torch.randn(n_e, d, device="cuda", requires_grad=True)
    sbf1 = torch.randn(g.idx_jj_pair.shape[0], d, device="cuda", requires_grad=True)
    sbf2 = torch.randn(g.idx_kj.shape[0], d, device="cuda", requires_grad=True)
    c = lambda t: t.detach().double().cpu().requires_grad_(True)
    x64, rbf64, s164, s264 = c(x), c(rbf), c(sbf1), c(sbf2)
    dev = lambda t: t.cuda()
    if simple:
        outs = layer(x, rbf, sbf1, dev(g.idx_jj_pair), dev(g.idx_ji_pair), dev(ei))
        refs = O.local_layer(sd, "L", x64, rbf64, None, s164, None, None, g.idx_jj_pair, g.idx_ji_pair, ei, two_hop=False)
        _check(outs, refs, [x, rbf, sbf1], [x64, rbf64, s164], layer, sd, "L")
    else:
        outs = layer(x, rbf, sbf2, sbf1, dev(g.idx_kj), dev(g.idx_ji), dev(g.idx_jj_pair), dev(g.idx_ji_pair), dev(ei))
        refs = O.local_layer(sd, "L", x64, rbf64, s264, s164, g.idx_kj, g.idx_ji, g.idx_jj_pair, g.idx_ji_pair, ei)
        _check(outs, refs, [x, rbf, sbf2, sbf1], [x64, rbf64, s264, s164], layer, sd, "L")


def test_layers_run_without_grad_too():
    from pamnet_b200 import Config
    from pamnet_b200.layers import Global_MessagePassing
    b, g = _graph(2)
    layer = Global_MessagePassing(Config("QM9", 32, 1, 5.0, 5.0)).cuda()
    x = torch.randn(b.x.shape[0], 32, device="cuda")
    ea = torch.randn(g.edge_index_g.shape[1], 32, device="cuda")
    with torch.no_grad():
        a = layer(x, ea, g.edge_index_g.cuda())
    bb = layer(x.requires_grad_(True), ea, g.edge_index_g.cuda())
    for u, v in zip(a, bb):
        assert rel_err(u, v.detach()) < 1e-6
