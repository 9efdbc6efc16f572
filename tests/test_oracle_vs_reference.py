"""CPU, build container only: the oracle restatement against the reference's models.py imported
VERBATIM from /root/reference (skipped wherever the reference tree is absent, e.g. the GPU box)."""
import pytest
import torch

from oracle import ref_shim, pamnet_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def test_parameter_counts(ref):
    """main_qm9.py:26-27 count: 3 581 100 (PAMNet) / 3 573 292 (PAMNet_s) at dim=128, L=6; RNA 11 714."""
    cfg = ref.Config("QM9", 128, 6, 5.0, 5.0)
    n = lambda simple, c: sum(int(torch.tensor(s).prod()) for _, s in O.param_shapes(c, simple))
    assert n(False, cfg) == 3581100 and n(True, cfg) == 3573292
    assert n(False, ref.Config("rna_native", 16, 1, 2.6, 20.0)) == 11714


@pytest.mark.parametrize("simple", [False, True])
def test_qm9_forward_backward_matches_verbatim_reference(ref, simple):
    from pamnet_b200.data import synthetic_qm9_batch
    torch.manual_seed(3)
    cfg = ref.Config("QM9", 16, 2, 5.0, 5.0)
    model = (ref.PAMNet_s if simple else ref.PAMNet)(cfg).double()
    b = synthetic_qm9_batch(6, seed=5)
    b64 = type(b)(**{**b.__dict__, "pos": b.pos.double()})
    out = model(b64)
    (out - b.y.double()).abs().mean().backward()
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    mine = O.forward(leaves, cfg, b, simple=simple)
    (mine - b.y.double()).abs().mean().backward()
    assert rel_err(mine, out) < 1e-6
    for k, p in model.named_parameters():
        if p.grad is None:
            assert leaves[k].grad is None
        else:
            assert rel_err(leaves[k].grad, p.grad) < 2e-6, k


def test_edge_permutation_invariance(ref):
    """licence to CSR-sort edges internally."""
    from pamnet_b200.data import synthetic_qm9_batch
    cfg = ref.Config("QM9", 16, 1, 5.0, 5.0)
    sd = O.init_state_dict(cfg, seed=1, dtype=torch.float64)
    b = synthetic_qm9_batch(3, seed=2)
    perm = torch.randperm(b.edge_index.shape[1], generator=torch.Generator().manual_seed(0))
    b2 = type(b)(**{**b.__dict__, "edge_index": b.edge_index[:, perm]})
    assert rel_err(O.forward(sd, cfg, b2), O.forward(sd, cfg, b)) < 1e-12
