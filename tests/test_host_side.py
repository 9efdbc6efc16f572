"""CPU: host logic -- C-ABI library loads and exports every declared symbol, parameter layout == reference
state_dict, module surface (keys, counts, flattening), loud failure without CUDA.  No GPU compute."""
import os
import re

import pytest
import torch

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pamnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pamnet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pamnet_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pamnet_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.pamnet_abi_version() == 1


def test_bad_config_is_an_error_not_a_crash():
    from pamnet_b200 import _lib
    lib = _lib.load()
    assert lib.pamnet_param_count(_lib.Config(0, 100, 6, 0, 0, 5.0, 5.0)) < 0       # unsupported dim
    assert b"dim" in lib.pamnet_last_error()
    assert lib.pamnet_param_count(_lib.Config(2, 16, 1, 0, 1, 2.6, 20.0)) < 0        # PAMNet_s is QM9-only


@pytest.mark.parametrize("name,simple", [("qm9_small_pamnet", False), ("qm9_small_pamnet_s", True),
                                         ("pdbbind_small", False), ("rna_native", False)])
def test_param_layout_matches_reference_state_dict(name, simple):
    from pamnet_b200 import _lib, Config, PAMNet, PAMNet_s
    gold = load_golden(name)
    model = (PAMNet_s if simple else PAMNet)(Config(**gold["config"]))
    keys = list(model.state_dict().keys())
    if name == "rna_native":            # shipped checkpoint predates the final registration order
        assert sorted(keys) == sorted(gold["state_dict"].keys())
    else:
        assert keys == list(gold["state_dict"].keys())
    for k, v in gold["state_dict"].items():
        assert tuple(model.state_dict()[k].shape) == tuple(v.shape), k
    assert all(off % 32 == 0 for off in model._offsets)                 # 128 B aligned tensors
    assert model._aliased()
    res = model.load_state_dict(gold["state_dict"])
    assert not res.missing_keys and not res.unexpected_keys
    assert model._aliased()                                             # load_state_dict copies in place
    for k, v in gold["state_dict"].items():
        assert torch.equal(model.state_dict()[k], v)
    flat = model._flat
    for (n, p), off in zip(model._param_list, model._offsets):
        assert torch.equal(flat[off:off + p.numel()].view(p.shape), p.detach())


def test_parameter_counts():
    """main_qm9.py:26-27 at the README config (README.md:95): 3 581 100 / 3 573 292; RNA checkpoint 11 714."""
    from pamnet_b200 import Config, PAMNet, PAMNet_s
    n = lambda m: sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert n(PAMNet(Config("QM9", 128, 6, 5.0, 5.0))) == 3581100
    assert n(PAMNet_s(Config("QM9", 128, 6, 5.0, 5.0))) == 3573292
    m = PAMNet(Config("rna_native", 16, 1, 2.6, 20.0, "target_to_source"))
    assert n(m) == 11714 and len(list(m.parameters())) == 74


def test_reflatten_after_param_data_swap():
    """utils/ema.py:27,32 replace param.data wholesale; the model must notice and re-pack."""
    from pamnet_b200 import Config, PAMNet
    m = PAMNet(Config("QM9", 16, 1, 5.0, 5.0))
    p = dict(m.named_parameters())["global_layer.0.mlp_x1.0.0.weight"]
    p.data = torch.full_like(p.data, 0.5)
    assert not m._aliased()
    m._flatten()
    assert m._aliased() and float(p.detach().mean()) == 0.5
    off = m._offsets[[n for n, _ in m._param_list].index("global_layer.0.mlp_x1.0.0.weight")]
    assert float(m._flat[off]) == 0.5


def test_errors_match_the_reference():
    from pamnet_b200 import Config, PAMNet, PAMNet_s, Batch
    b = Batch(x=torch.zeros(3), pos=torch.zeros(3, 3), edge_index=torch.zeros(2, 0, dtype=torch.long),
              batch=torch.zeros(3, dtype=torch.long), y=torch.zeros(1))
    with pytest.raises(ValueError, match="Invalid dataset"):           # models.py:160
        PAMNet(Config("ZINC", 16, 1, 5.0, 5.0))(b)
    with pytest.raises(ValueError, match="only for QM9"):              # models.py:287
        PAMNet_s(Config("PDBbind", 16, 1, 5.0, 5.0))(b)
    with pytest.raises(ValueError):
        PAMNet(Config("QM9", 100, 1, 5.0, 5.0))                        # dim the kernels do not implement


def test_cpu_tensors_fail_loudly_no_fallback():
    from pamnet_b200 import Config, PAMNet, PamnetError, ops
    from pamnet_b200.data import synthetic_qm9_batch
    with pytest.raises(PamnetError, match="CUDA"):
        PAMNet(Config("QM9", 16, 1, 5.0, 5.0))(synthetic_qm9_batch(2))
    with pytest.raises(PamnetError):
        ops.radius_graph(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long), 1.0)
    with pytest.raises(PamnetError, match="CUDA"):                      # the opt-in plan prefetch has no CPU path either
        PAMNet(Config("QM9", 16, 1, 5.0, 5.0)).prefetch(synthetic_qm9_batch(2))


def test_synthetic_generator_contract():
    """12-28 atoms, min separation 0.9, degree <= 4, symmetric sorted bonds."""
    from pamnet_b200.data import synthetic_qm9_batch
    b = synthetic_qm9_batch(32, seed=0)
    assert b.num_graphs == 32 and b.x.dtype == torch.float32 and b.edge_index.dtype == torch.int64
    counts = torch.bincount(b.batch)
    assert int(counts.min()) >= 12 and int(counts.max()) <= 28
    assert bool((b.batch[1:] >= b.batch[:-1]).all())
    r, c = b.edge_index
    assert bool((b.batch[r] == b.batch[c]).all()) and int(torch.bincount(r).max()) <= 4
    fwd = set(zip(r.tolist(), c.tolist()))
    assert all((j, i) in fwd for i, j in fwd)
    d = (b.pos[r] - b.pos[c]).norm(dim=1)
    assert float(d.min()) >= 0.9 and float(d.max()) < 1.7
    b2 = synthetic_qm9_batch(32, seed=0)
    assert torch.equal(b.pos, b2.pos) and torch.equal(b.edge_index, b2.edge_index)


def test_fused_optimizer_refuses_cpu_parameters_and_foreign_modules():
    """pamnet_b200.FusedAdamEMA has no CPU path and only drives this package's flat-parameter modules."""
    from pamnet_b200 import Config, PAMNet, FusedAdamEMA, PamnetError
    model = PAMNet(Config("QM9", 16, 1, 5.0, 5.0))
    with pytest.raises(PamnetError):
        FusedAdamEMA(model, lr=1e-3)
    with pytest.raises(TypeError):
        FusedAdamEMA(torch.nn.Linear(4, 4), lr=1e-3)


def test_optimizer_step_argument_checks_need_no_gpu():
    """The C entry point validates sizes before touching memory (n must be a multiple of 4, step >= 1)."""
    import ctypes
    from pamnet_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 8)()
    p = ctypes.addressof(buf)
    rc = lib.pamnet_optimizer_step(p, p, p, p, None, 6, None, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0, 0.0, 0, None, None)
    assert rc < 0 and b"multiple of 4" in lib.pamnet_last_error()
    rc = lib.pamnet_optimizer_step(p, p, p, p, None, 8, None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.0, 0.0, 0, None, None)
    assert rc < 0


def test_plan_build_argument_checks_need_no_gpu():
    import ctypes
    from pamnet_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(0, 16, 1, 0, 0, 5.0, 5.0)
    assert lib.pamnet_plan_build_scratch_bytes(cfg, 100, 200, 1000) > 0
    assert lib.pamnet_prepared_weights_bytes(cfg) > 0
    need = (ctypes.c_int64 * 4)()
    sz = _lib.Sizes(0, 0, 0, 0, 0, 0)
    rc = lib.pamnet_plan_build(cfg, None, None, 10, 1, None, 0, 32, None, 0, None, 0, None, 0, None, 0, None, 0, sz, need, None)
    assert rc < 0 and b"null pointer" in lib.pamnet_last_error()
