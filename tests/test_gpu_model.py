"""GPU: the fused CUDA model path (C ABI behind pamnet_b200.PAMNet) against the golden vectors produced by
the verbatim reference and against the oracle at BASELINE.json's full size.

Tolerance (SURVEY.md 8(c) ladder): fp64 golden / oracle is truth; accept err <= max(1e-5 rel, 2 * err of the
fp32 reference), per tensor, relative to max|truth|.  Integer outputs bit-exact."""
import pytest
import torch

from tests.helpers import load_golden, cfg_of, batch_of, rel_err, ladder_ok, oracle_step, elementwise_ok

pytestmark = pytest.mark.gpu


def _model(gold, simple=False):
    from pamnet_b200 import Config, PAMNet, PAMNet_s
    m = (PAMNet_s if simple else PAMNet)(Config(**gold["config"]))
    m.load_state_dict(gold["state_dict"])
    return m.cuda()


def _step(model, batch, loss="l1"):
    b = batch.to("cuda")
    model.zero_grad()
    out = model(b)
    l = (out - b.y).abs().mean() if loss == "l1" else ((out - b.y) ** 2).mean()
    l.backward()
    torch.cuda.synchronize()
    return out.detach().cpu(), l.detach().cpu(), {k: (p.grad.cpu() if p.grad is not None else None)
                                                   for k, p in model.named_parameters()}


@pytest.mark.parametrize("name,simple,loss", [("qm9_small_pamnet", False, "l1"), ("qm9_small_pamnet_s", True, "l1"),
                                              ("pdbbind_small", False, "mse"), ("rna_native", False, "l1")])
def test_golden_forward_backward(name, simple, loss):
    gold = load_golden(name)
    model = _model(gold, simple)
    out, l, grads = _step(model, batch_of(gold), loss)
    ok, e_new, e_ref = ladder_ok(out, gold["out_f32"], gold["out_f64"])
    assert ok, ("out", e_new, e_ref)
    assert rel_err(out, gold["out_f32"]) < 1e-5
    bad = []
    for k, ref64 in gold["grads_f64"].items():
        if ref64 is None:
            assert grads[k] is None, k
            continue
        ok, e_new, e_ref = ladder_ok(grads[k], gold["grads_f32"][k], ref64)
        if not ok:
            bad.append((k, e_new, e_ref))
        ok, ratio = elementwise_ok(grads[k], gold["grads_f32"][k], ref64)     # no small-magnitude row hides behind the max-norm
        if not ok:
            bad.append((k, "elementwise", ratio))
    assert not bad, bad[:10]


def test_c4_real_fixture_first_8_natives():
    """BASELINE.json configs[3] on its OWN data: the first 8 graphs of the reference's rna_native
    fixture (841-3 771 atoms, 15 816 in total) in one batch, shipped checkpoint, forward + L1 + backward against the
    verbatim reference's fp32 / fp64 results (tests/golden/rna_c4.pt, make_golden.py:make_rna_c4)."""
    from pamnet_b200.data import Batch
    gold = load_golden("rna_c4")
    model = _model(gold)
    sizes = gold["sizes"]
    assert sizes == [1384, 1176, 1246, 3215, 932, 3771, 841, 3251]
    batch = Batch(x=gold["x"], batch=torch.repeat_interleave(torch.arange(8), torch.tensor(sizes)), y=gold["y"])
    out, l, grads = _step(model, batch)
    n = sum(sizes)
    plan = model.last_plan
    assert plan.sizes.n_nodes == n and plan.sizes.n_edges_g == 49 * n            # kNN-50 minus self, nothing beyond 20 A
    assert (plan.sizes.n_edges_l, plan.sizes.n_t2, plan.sizes.n_t1) == (86160, 432574, 518734)
    ok, e_new, e_ref = ladder_ok(out, gold["out_f32"], gold["out_f64"])
    assert ok, ("out", e_new, e_ref)
    assert rel_err(out, gold["out_f32"]) < 1e-5
    bad = []
    for k, ref64 in gold["grads_f64"].items():
        if ref64 is None:
            assert grads[k] is None, k
            continue
        ok, e_new, e_ref = ladder_ok(grads[k], gold["grads_f32"][k], ref64)
        if not ok:
            bad.append((k, e_new, e_ref))
    assert not bad, bad[:10]


def test_golden_graph_is_bit_exact():
    gold = load_golden("qm9_small_pamnet")
    model = _model(gold)
    with torch.no_grad():
        model(batch_of(gold).to("cuda"))
    plan = model.last_plan
    assert torch.equal(plan.edge_index_g.cpu(), gold["graph"]["edge_index_g"])
    assert torch.equal(plan.edge_index_l.cpu(), gold["graph"]["edge_index_l"])
    assert plan.sizes.n_t2 == gold["graph"]["idx_kj"].numel() and plan.sizes.n_t1 == gold["graph"]["idx_jj_pair"].numel()


def test_rna_native_scores():
    """Shipped checkpoint x shipped native structures -> the reference's scores (README.md:107-109 workflow)."""
    gold = load_golden("rna_native")
    model = _model(gold).eval()
    from pamnet_b200.data import Batch
    for g in gold["shipped"]:
        x = gold["inputs"][g].cuda()
        b = Batch(x=x, batch=torch.zeros(x.shape[0], dtype=torch.long, device="cuda"), y=torch.zeros(1, device="cuda"))
        with torch.no_grad():
            out = model(b)
        assert rel_err(out, gold["scores_f64"][g:g + 1]) < 1e-5


@pytest.mark.parametrize("n_graphs,dim,n_layer", [(32, 128, 6), (5, 64, 1), (3, 16, 2)])
def test_full_size_vs_oracle(n_graphs, dim, n_layer):
    """BASELINE.json configs[1] (QM9 dim=128 L=6 bs=32) and smaller shapes: CUDA vs oracle (fp32 and fp64)."""
    import types
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    from oracle import pamnet_oracle as O
    cfg = types.SimpleNamespace(dataset="QM9", dim=dim, n_layer=n_layer, cutoff_l=5.0, cutoff_g=5.0,
                                flow="source_to_target")
    sd = O.init_state_dict(cfg, seed=0)
    b = synthetic_qm9_batch(n_graphs, seed=0)
    model = PAMNet(Config(**vars(cfg)))
    model.load_state_dict(sd)
    out, l, grads = _step(model.cuda(), b)
    o32, _, g32 = oracle_step(sd, cfg, b, dtype=torch.float32)
    o64, _, g64 = oracle_step(sd, cfg, b, dtype=torch.float64)
    ok, e_new, e_ref = ladder_ok(out, o32, o64)
    assert ok, (e_new, e_ref)
    bad = []
    for k, ref64 in g64.items():
        if ref64 is None:
            continue
        ok, e_new, e_ref = ladder_ok(grads[k], g32[k], ref64)
        if not ok:
            bad.append((k, e_new, e_ref))
    assert not bad, bad[:10]


def test_edge_cases_isolated_atoms_and_single_atom_graph():
    """A graph with one atom (no edges at all) and an atom with no bonds inside a molecule."""
    import types
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch, Batch
    from oracle import pamnet_oracle as O
    b = synthetic_qm9_batch(2, seed=4)
    n = b.pos.shape[0]
    keep = (b.edge_index[0] != 1) & (b.edge_index[1] != 1)          # atom 1 loses its bonds
    x = torch.cat([b.x, torch.tensor([2.0])])
    pos = torch.cat([b.pos, torch.tensor([[50.0, 50.0, 50.0]])])
    batch = torch.cat([b.batch, torch.tensor([2])])
    bb = Batch(x=x, pos=pos, edge_index=b.edge_index[:, keep], batch=batch, y=torch.tensor([0.1, -0.2, 0.3]))
    cfg = types.SimpleNamespace(dataset="QM9", dim=32, n_layer=2, cutoff_l=5.0, cutoff_g=5.0, flow="source_to_target")
    sd = O.init_state_dict(cfg, seed=1)
    model = PAMNet(Config(**vars(cfg)))
    model.load_state_dict(sd)
    out, l, grads = _step(model.cuda(), bb)
    o64, _, g64 = oracle_step(sd, cfg, bb, dtype=torch.float64)
    assert rel_err(out, o64) < 1e-5
    for k, ref in g64.items():
        if ref is not None:
            assert rel_err(grads[k], ref) < 2e-5, k


def test_accumulates_like_autograd_and_is_deterministic():
    gold = load_golden("qm9_small_pamnet")
    model = _model(gold)
    b = batch_of(gold).to("cuda")
    model.zero_grad()
    (model(b) - b.y).abs().mean().backward()
    g1 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    (model(b) - b.y).abs().mean().backward()                       # no zero_grad: grads must double
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert rel_err(p.grad, 2 * g1[k]) < 1e-6, k
    model.zero_grad()
    (model(b) - b.y).abs().mean().backward()
    same = all(torch.equal(p.grad, g1[k]) for k, p in model.named_parameters() if p.grad is not None)
    assert same or max(rel_err(p.grad, g1[k]) for k, p in model.named_parameters() if p.grad is not None) < 1e-6


def test_cpu_input_fails_loudly():
    from pamnet_b200 import PamnetError
    gold = load_golden("qm9_small_pamnet")
    model = _model(gold)
    with pytest.raises((PamnetError, RuntimeError)):
        model(batch_of(gold))


# ---- BASELINE.json configs[2] / configs[3] at full size: size-independent properties ---------------------------------
def _split_batch(b, lo, hi):
    """Graphs lo..hi-1 of a collated batch as their own batch (node / edge indices re-based)."""
    from pamnet_b200.data import Batch
    nodes = (b.batch >= lo) & (b.batch < hi)
    first = int(nodes.nonzero()[0])
    f = dict(x=b.x[nodes], batch=b.batch[nodes] - lo, y=b.y[lo:hi])
    if getattr(b, "pos", None) is not None:
        f["pos"] = b.pos[nodes]
    if getattr(b, "edge_index", None) is not None:
        e = nodes[b.edge_index[0]]
        f["edge_index"] = b.edge_index[:, e] - first
    return Batch(**f)


def test_c3_batch256_equals_its_32_molecule_shards():
    """configs[2] size (QM9 dim=128 L=6 bs=256; fp32 -- the bf16 node-MLP variant is not built): molecules are
    independent, so the batch-256 outputs are the outputs of its eight 32-molecule shards and the
    gradient of the summed loss is the sum of the shard gradients."""
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    torch.manual_seed(0)
    model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
    big = synthetic_qm9_batch(256, seed=0)
    model.zero_grad()
    out = model(big.to("cuda"))
    (out - big.y.cuda()).abs().sum().backward()
    g_big = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    outs = []
    for s in range(8):
        sb = _split_batch(big, 32 * s, 32 * (s + 1)).to("cuda")
        o = model(sb)
        (o - sb.y).abs().sum().backward()            # accumulates over the shards
        outs.append(o.detach())
    assert rel_err(out.detach(), torch.cat(outs)) < 2e-6
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert rel_err(g_big[k], p.grad) < 2e-5, k


def test_c4_rna_batch8_large_graphs():
    """configs[3] size (rna dim=16 L=1, 8 graphs of ~2k atoms, target_to_source, kNN-50): batched == per graph, every
    node has exactly 49 incoming global messages, and a 3-graph slice matches the fp64 oracle forward + backward."""
    import types
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_rna_batch
    from oracle import pamnet_oracle as O
    cfg = types.SimpleNamespace(dataset="rna_native", dim=16, n_layer=1, cutoff_l=2.6, cutoff_g=20.0, flow="target_to_source")
    sd = O.init_state_dict(cfg, seed=0)
    model = PAMNet(Config(**vars(cfg)))
    model.load_state_dict(sd)
    model = model.cuda()
    # ~2k-atom graphs without the generator's quadratic rejection loop: five translated copies of a 300-500 atom chain
    from pamnet_b200.data import Batch
    base = synthetic_rna_batch(8, seed=0, min_atoms=300, max_atoms=500)
    xs, bs = [], []
    for g in range(8):
        xg = base.x[base.batch == g]
        for c in range(5):
            xs.append(xg + torch.tensor([37.0 * c, 11.0 * c, 0.0, 0.0]))
            bs.append(torch.full((xg.shape[0],), g, dtype=torch.long))
    big = Batch(x=torch.cat(xs), batch=torch.cat(bs), y=base.y)
    with torch.no_grad():
        out = model(big.to("cuda"))
        plan = model.last_plan
        n = big.x.shape[0]
        assert plan.sizes.n_edges_g == 49 * n                       # E_g = 49 N exactly
        assert torch.equal(torch.bincount(plan.edge_index_g[0].cpu(), minlength=n), torch.full((n,), 49))
        per = torch.cat([model(_split_batch(big, g, g + 1).to("cuda")) for g in range(8)])
    assert rel_err(out, per) < 2e-6
    small = synthetic_rna_batch(3, seed=1, min_atoms=300, max_atoms=500)
    o, l, grads = _step(model, small)
    o64, _, g64 = oracle_step(sd, cfg, small, dtype=torch.float64)
    o32, _, g32 = oracle_step(sd, cfg, small, dtype=torch.float32)
    ok, e_new, e_ref = ladder_ok(o, o32, o64)
    assert ok, (e_new, e_ref)
    for k, ref64 in g64.items():
        if ref64 is not None:
            ok, e_new, e_ref = ladder_ok(grads[k], g32[k], ref64)
            assert ok, (k, e_new, e_ref)


# ---- alternate execution paths must give the same numbers ------------------------------------------------------------
def _run_subprocess_step(env):
    """forward + backward of the QM9 golden model in a fresh process with `env` set; returns (out, one gradient)."""
    import os, subprocess, sys, tempfile
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from tests.helpers import load_golden, batch_of
from pamnet_b200 import Config, PAMNet
gold = load_golden("qm9_small_pamnet")
m = PAMNet(Config(**gold["config"])); m.load_state_dict(gold["state_dict"]); m = m.cuda()
b = batch_of(gold).to("cuda")
out = m(b); (out - b.y).abs().mean().backward(); torch.cuda.synchronize()
g = dict(m.named_parameters())["global_layer.0.mlp_m.0.0.weight"].grad
torch.save({"out": out.detach().cpu(), "g": g.cpu(), "eg": m.last_plan.edge_index_g.cpu(), "el": m.last_plan.edge_index_l.cpu()}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        subprocess.run([sys.executable, "-c", code, f.name], check=True, env={**os.environ, **env}, timeout=300)
        return torch.load(f.name)


_BASE_STEP = None


@pytest.mark.parametrize("env", [{"PAMNET_STREAMS": "1"}, {"PAMNET_PDL": "0"}, {"PAMNET_PLAN": "stepwise"},
                                 {"PAMNET_GATHER": "atomic"}, {"PAMNET_GEMM": "ffma"}, {"PAMNET_CHAIN_FUSE": "0"},
                                 {"PAMNET_PREP": "1"}, {"PAMNET_FWD_SPLIT": "1"}, {"PAMNET_GEMM_SMALL": "0"},
                                 {"PAMNET_SBF_FUSED": "0"}, {"PAMNET_GEMM": "tc1"}, {"PAMNET_CHAIN": "ffma"},
                                 {"PAMNET_FRONT": "generic"}])
def test_alternate_paths_agree(env):
    """Single-stream schedule, no programmatic dependent launch, the step-by-step front end, atomic projection
    gradients, the FFMA GEMM, the unfused chain prologue, ... (the golden model has dim = 32, so the
    skinny-GEMM and fused spherical-basis paths are the defaults here): same graph bit for bit, same numbers to fp32
    rounding."""
    global _BASE_STEP
    if _BASE_STEP is None:
        _BASE_STEP = _run_subprocess_step({})
    base = _BASE_STEP
    alt = _run_subprocess_step(env)
    assert torch.equal(base["eg"], alt["eg"]) and torch.equal(base["el"], alt["el"])
    assert rel_err(alt["out"], base["out"]) < 2e-6
    assert rel_err(alt["g"], base["g"]) < 1e-5


def test_plan_build_grows_capacities():
    """pamnet_plan_build with deliberately tiny capacities: returns 1 with the needed sizes, the retry succeeds and the
    plan equals the one built step by step."""
    import os
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch, synthetic_rna_batch
    for kind, cfgargs, batch in [("QM9", ("QM9", 32, 1, 5.0, 5.0), synthetic_qm9_batch(5, seed=2)),
                                 ("rna", ("rna_native", 16, 1, 2.6, 20.0, "target_to_source"), synthetic_rna_batch(2, seed=2))]:
        model = PAMNet(Config(*cfgargs)).cuda()
        b = batch.to("cuda")
        model._plan_caps = {"eg": 8, "el": 8, "base": 64, "trip": 64}
        with torch.no_grad():
            out = model(b)
        fused = model.last_plan
        assert model._plan_caps["eg"] >= fused.sizes.n_edges_g
        os.environ["PAMNET_PLAN"] = "stepwise"
        try:
            with torch.no_grad():
                out2 = model(b)
        finally:
            del os.environ["PAMNET_PLAN"]
        step = model.last_plan
        assert torch.equal(fused.edge_index_g, step.edge_index_g) and torch.equal(fused.edge_index_l, step.edge_index_l)
        assert (fused.sizes.n_t2, fused.sizes.n_t1) == (step.sizes.n_t2, step.sizes.n_t1)
        assert torch.equal(out, out2), kind


def test_backends_agree_at_dim_128():
    """The golden model has dim 32 (FFMA chain, skinny GEMMs); the tensor-core node chain (chain_mma.cu) and the TMA
    GEMM (gemm_tc2.cu) only run at dim >= 64.  Same step at dim 128 through {tensor-core chain + TMA GEMM} (default),
    the round-1 GEMM, and the all-FFMA path: equal to fp32 rounding."""
    import os, subprocess, sys, tempfile
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
torch.manual_seed(0)
m = PAMNet(Config("QM9", 128, 2, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(16, seed=3).to("cuda")
out = m(b); (out - b.y).abs().mean().backward(); torch.cuda.synchronize()
torch.save({"out": out.detach().cpu(), "g": torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu()}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, env in [("default", {}), ("tc1", {"PAMNET_GEMM": "tc1"}), ("ffma", {"PAMNET_GEMM": "ffma", "PAMNET_CHAIN": "ffma"})]:
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env={**os.environ, **env}, timeout=300)
            res[name] = torch.load(f.name)
    for name in ("default", "tc1"):
        assert rel_err(res[name]["out"], res["ffma"]["out"]) < 2e-6, name
        assert rel_err(res[name]["g"], res["ffma"]["g"]) < 1e-5, name


def test_reduced_precision_node_mlp_rung():
    """BASELINE.json configs[2] names a reduced-precision tensor-core node-MLP path.  PAMNET_NODE_MLP=tf32 runs the node
    chains as ONE tensor-core pass on tf32-truncated operands (10-bit mantissa, >= bf16's 7) with fp32 accumulation;
    everything else stays fp32-accurate.  Its own rung, stated against the fp64 oracle and NOT part of the 1e-5 ladder:
    outputs within 5e-3, every gradient tensor within 5e-2 of max|fp64| (measured on a B200: 2e-3 / 1e-2)."""
    import os, subprocess, sys, tempfile, types
    from pamnet_b200.data import synthetic_qm9_batch
    from oracle import pamnet_oracle as O
    cfg = types.SimpleNamespace(dataset="QM9", dim=128, n_layer=2, cutoff_l=5.0, cutoff_g=5.0, flow="source_to_target")
    sd = O.init_state_dict(cfg, seed=0)
    b = synthetic_qm9_batch(8, seed=5)
    o64, _, g64 = oracle_step(sd, cfg, b, dtype=torch.float64)
    code = r'''
import sys, types, torch
sys.path.insert(0, %r)
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
from oracle import pamnet_oracle as O
cfg = types.SimpleNamespace(dataset="QM9", dim=128, n_layer=2, cutoff_l=5.0, cutoff_g=5.0, flow="source_to_target")
m = PAMNet(Config(**vars(cfg))); m.load_state_dict(O.init_state_dict(cfg, seed=0)); m = m.cuda()
b = synthetic_qm9_batch(8, seed=5).to("cuda")
out = m(b); (out - b.y).abs().mean().backward(); torch.cuda.synchronize()
torch.save({"out": out.detach().cpu(), "g": {k: (p.grad.cpu() if p.grad is not None else None) for k, p in m.named_parameters()}}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        subprocess.run([sys.executable, "-c", code, f.name], check=True, env={**os.environ, "PAMNET_NODE_MLP": "tf32"}, timeout=300)
        res = torch.load(f.name)
    e_out = rel_err(res["out"], o64)
    assert 1e-7 < e_out < 5e-3, e_out           # really reduced precision, and within its rung
    worst = max(rel_err(res["g"][k], v) for k, v in g64.items() if v is not None)
    assert worst < 5e-2, worst


def test_second_backward_raises_clearly():
    """The step's workspace is released after backward: a second backward through the same forward is an explicit error
    (not an AttributeError), as documented in _PAMNetBase."""
    gold = load_golden("qm9_small_pamnet")
    model = _model(gold)
    b = batch_of(gold).to("cuda")
    loss = (model(b) - b.y).abs().mean()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


@pytest.mark.gpu
def test_changing_batch_shapes_reuse_nothing_stale():
    """Every batch of an epoch has its own atom / edge / triplet counts, so activation operands change address and shape
    from step to step (by-value tensor maps) while the weights stay put (device tensor-map table): A, B, C, A again must
    give A's outputs and gradients again (to summation-order noise: split reductions accumulate with atomics)."""
    import torch
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    model = PAMNet(Config(dataset="QM9", dim=128, n_layer=2, cutoff_l=5.0, cutoff_g=5.0)).to(dev)
    batches = [synthetic_qm9_batch(n, seed=s).to(dev) for n, s in ((5, 1), (17, 2), (32, 3))]

    def run(b):
        model.zero_grad()
        out = model(b)
        (out - b.y).abs().mean().backward()
        return out.detach().clone(), [p.grad.detach().clone() for p in model.parameters() if p.grad is not None]

    first = run(batches[0])
    for b in batches[1:]:
        run(b)
    again = run(batches[0])
    assert rel_err(again[0], first[0]) < 1e-6
    assert len(first[1]) == len(again[1]) > 0
    for g0, g1 in zip(first[1], again[1]):
        assert rel_err(g1, g0) < 1e-5
