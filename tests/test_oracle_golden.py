"""CPU: the oracle restatement against the golden vectors produced by the verbatim reference
(tests/golden/make_golden.py).  Tolerances: fp64 rung 1e-6 rel (normaliser constants are fp32-rounded
under the container's numpy 2, see oracle/pamnet_oracle.py), fp32 rung 1e-5 rel; integers bit-exact."""
import pytest
import torch

from tests.helpers import load_golden, cfg_of, batch_of, rel_err, oracle_step, ladder_ok
from oracle import pamnet_oracle as O, graph_ops as G

CASES = [("qm9_small_pamnet", False, "l1"), ("qm9_small_pamnet_s", True, "l1"),
         ("pdbbind_small", False, "mse"), ("rna_native", False, "l1")]


@pytest.mark.parametrize("name,simple,loss", CASES)
def test_param_table_matches_reference_state_dict(name, simple, loss):
    gold = load_golden(name)
    table = O.param_shapes(cfg_of(gold), simple)
    if name == "rna_native":     # shipped checkpoint was saved by an older module (sbf2 before sbf1)
        assert sorted(k for k, _ in table) == sorted(gold["state_dict"].keys())
    else:
        assert [k for k, _ in table] == list(gold["state_dict"].keys())
    for k, shape in table:
        assert tuple(gold["state_dict"][k].shape) == shape, k


@pytest.mark.parametrize("name,simple,loss", CASES)
def test_oracle_forward_backward_vs_golden(name, simple, loss):
    """fp64 rung: tight against the reference's own fp64 run.  fp32 rung: precision ladder."""
    gold = load_golden(name)
    cfg, b = cfg_of(gold), batch_of(gold)
    out64, l64, g64 = oracle_step(gold["state_dict"], cfg, b, simple, loss, torch.float64)
    assert rel_err(out64, gold["out_f64"]) < 1e-6 and rel_err(l64, gold["loss_f64"]) < 1e-6
    out32, l32, g32 = oracle_step(gold["state_dict"], cfg, b, simple, loss, torch.float32)
    ok, e_new, e_ref = ladder_ok(out32, gold["out_f32"], gold["out_f64"])
    assert ok, (e_new, e_ref)
    for k, ref64 in gold["grads_f64"].items():
        if ref64 is None:
            assert g32[k] is None or float(g32[k].abs().max()) == 0.0, k
            continue
        assert rel_err(g64[k], ref64) < 2e-6, k
        ok, e_new, e_ref = ladder_ok(g32[k], gold["grads_f32"][k], ref64)
        assert ok, (k, e_new, e_ref)


def test_graph_vectors_bit_exact():
    gold = load_golden("qm9_small_pamnet")
    b, gv = batch_of(gold), gold["graph"]
    row, col = G.radius_pairs(b.pos, b.pos, 5.0, b.batch, b.batch, 1000)
    assert torch.equal(row, gv["radius_row"]) and torch.equal(col, gv["radius_col"])
    eg = G.drop_self_loops(torch.stack([row, col]))
    assert torch.equal(eg, gv["edge_index_g"])
    el = G.drop_self_loops(b.edge_index)
    names = ["idx_i", "idx_j", "idx_k", "idx_kj", "idx_ji", "idx_i_pair", "idx_j1_pair", "idx_j2_pair",
             "idx_jj_pair", "idx_ji_pair"]
    for n, v in zip(names, G.triplet_indices(el, b.pos.shape[0])):
        assert torch.equal(v, gv[n]), n
    # structural identities: T1 = T2 + E_l on a symmetric loop-free graph,
    # scatter keys are non-decreasing
    assert gv["idx_jj_pair"].numel() == gv["idx_kj"].numel() + el.shape[1]
    assert bool((gv["idx_ji"][1:] >= gv["idx_ji"][:-1]).all())


def test_rna_scores_regression():
    """The 21 native-structure scores of the shipped checkpoint (two structures shipped as inputs)."""
    gold = load_golden("rna_native")
    cfg = cfg_of(gold)
    from pamnet_b200.data import Batch
    for g in gold["shipped"]:
        x = gold["inputs"][g]
        b = Batch(x=x, batch=torch.zeros(x.shape[0], dtype=torch.long), y=torch.zeros(1))
        out = O.forward(gold["state_dict"], cfg, b)
        assert rel_err(out, gold["scores_f32"][g:g + 1]) < 1e-5
    assert abs(float(gold["scores_f32"][0]) - 2.665966) < 1e-5   # SURVEY.md section 4 vector


def test_sbf_constants_known_values():
    zeros, norm = O.sbf_constants()
    assert abs(zeros[1, 0] - 4.4934096) < 1e-6 and abs(zeros[6, 5] - 27.507868) < 1e-5
    assert abs(norm[0, 0] - 4.4428830) < 1e-6 and abs(norm[6, 0] - 16.68576) < 1e-4
    y = O.zonal_harmonic_coeffs()
    assert abs(y[0, 0] - 0.2820948) < 1e-7 and abs(y[6, 6] - 14.6844857) < 1e-6
