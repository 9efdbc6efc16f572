"""Per-molecule front end (PAMNET_FRONT=mol, csrc/front_mol.cuh) against the generic graph kernels on the GPU: every
plan array, both edge lists and the model output must be bit-identical; the generic plan is also checked against the numpy
restatement of the plan definition (tests/plan_ref.py).

Its integer logic is also covered on the CPU (tests/test_front_mol_host.py).  Green on a B200 since round 2
(gpurun_out/front_mol_parity.log); the per-molecule front end is the default for QM9-shaped batches, PAMNET_FRONT=generic
selects the generic kernels."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


def _plan(model, batch, front):
    old = os.environ.pop("PAMNET_FRONT", None)
    os.environ["PAMNET_FRONT"] = front or "generic"
    try:
        with torch.no_grad():
            out = model(batch)
    finally:
        os.environ.pop("PAMNET_FRONT", None)
        if old is not None:
            os.environ["PAMNET_FRONT"] = old
    plan = model.last_plan
    arrays = {k: v.cpu() for k, v in plan.arrays().items()}
    return out.cpu(), plan, arrays


@pytest.mark.parametrize("flow,simple", [("source_to_target", False), ("target_to_source", False), ("source_to_target", True)])
def test_mol_front_end_equals_generic(flow, simple):
    from pamnet_b200 import Config, PAMNet, PAMNet_s
    from pamnet_b200.data import synthetic_qm9_batch
    from pamnet_b200 import _lib
    lib = _lib.load()
    for n_graphs, seed in [(32, 0), (7, 3), (256, 1)]:
        torch.manual_seed(0)
        model = (PAMNet_s if simple else PAMNet)(Config("QM9", 32, 1, 5.0, 5.0, flow)).cuda()
        b = synthetic_qm9_batch(n_graphs, seed=seed).to("cuda")
        out_g, plan_g, arr_g = _plan(model, b, None)
        n0 = lib.pamnet_debug_launch_count()
        out_m, plan_m, arr_m = _plan(model, b, "mol")
        n_mol = lib.pamnet_debug_launch_count() - n0
        n0 = lib.pamnet_debug_launch_count()
        _plan(model, b, None)
        assert n_mol < lib.pamnet_debug_launch_count() - n0 - 10          # the per-molecule path really ran (2 launches for ~20)
        sg, sm = plan_g.sizes, plan_m.sizes
        assert (sg.n_edges_g, sg.n_edges_l, sg.n_t2, sg.n_t1) == (sm.n_edges_g, sm.n_edges_l, sm.n_t2, sm.n_t1)
        assert torch.equal(plan_g.edge_index_g, plan_m.edge_index_g)
        assert torch.equal(plan_g.edge_index_l, plan_m.edge_index_l)
        for k in arr_g:
            a, c = arr_g[k], arr_m[k]
            same = torch.equal(a.view(torch.int32), c.view(torch.int32))
            assert same, (k, n_graphs, flow)
        assert torch.equal(out_g, out_m)


def test_mol_front_end_filters_self_loops_and_falls_back():
    """Self loops in the bond list (a filtered list is written) and a shuffled bond list (generic fallback)."""
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    torch.manual_seed(0)
    model = PAMNet(Config("QM9", 32, 1, 5.0, 5.0)).cuda()
    b = synthetic_qm9_batch(6, seed=5)
    ei = b.edge_index
    loops = torch.tensor([[0, 3], [0, 3]])
    b.edge_index = torch.cat([loops, ei], dim=1)                        # atoms 0 and 3 are in molecule 0: stays grouped
    bc = b.to("cuda")
    out_g, plan_g, arr_g = _plan(model, bc, None)
    out_m, plan_m, arr_m = _plan(model, bc, "mol")
    assert plan_m.edge_index_l.shape[1] == ei.shape[1]
    assert torch.equal(plan_g.edge_index_l, plan_m.edge_index_l) and torch.equal(out_g, out_m)
    for k in arr_g:
        assert torch.equal(arr_g[k].view(torch.int32), arr_m[k].view(torch.int32)), k
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    b.edge_index = ei[:, perm]
    bc = b.to("cuda")
    out_g, plan_g, _ = _plan(model, bc, None)
    out_m, plan_m, _ = _plan(model, bc, "mol")
    assert torch.equal(plan_g.edge_index_l, plan_m.edge_index_l) and torch.equal(out_g, out_m)


def test_generic_plan_matches_definition():
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    from tests import plan_ref
    torch.manual_seed(0)
    for flow in ("source_to_target", "target_to_source"):
        model = PAMNet(Config("QM9", 32, 1, 5.0, 5.0, flow)).cuda()
        b = synthetic_qm9_batch(8, seed=6)
        _, plan, arr = _plan(model, b.to("cuda"), None)
        ref = plan_ref.build_plan(b.pos.numpy(), b.batch.numpy(), 8, plan.edge_index_g.cpu().numpy(),
                                  plan.edge_index_l.cpu().numpy(), 0 if flow == "target_to_source" else 1)
        for k, v in arr.items():
            if k == "t_angle":
                assert np.allclose(v.numpy(), ref[k], rtol=0, atol=2e-6)
            elif k.startswith("dist"):
                assert np.array_equal(v.numpy(), ref[k]), k
            else:
                assert np.array_equal(v.numpy().astype(np.int64), ref[k]), k


def test_prefetched_plan_gives_the_same_step():
    """model.prefetch(batch) builds the plan on a side stream; forward + backward from it equal the inline path bit for
    bit (gradients up to the summation order of the split-K atomics), a prefetch for another object is ignored, and it
    composes with the per-molecule front end."""
    from pamnet_b200 import Config, PAMNet
    from pamnet_b200.data import synthetic_qm9_batch
    torch.manual_seed(0)
    model = PAMNet(Config("QM9", 64, 2, 5.0, 5.0)).cuda()
    b = synthetic_qm9_batch(8, seed=7).to("cuda")
    other = synthetic_qm9_batch(3, seed=8).to("cuda")

    def run(batch, prefetch_of=None, front=None):
        os.environ["PAMNET_FRONT"] = front or "generic"
        try:
            for p in model.parameters():
                p.grad = None
            if prefetch_of is not None:
                model.prefetch(prefetch_of)
            out = model(batch)
            (out - batch.y).abs().mean().backward()
            torch.cuda.synchronize()
            return out.detach().clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).clone()
        finally:
            os.environ.pop("PAMNET_FRONT", None)

    def same_grads(a, c):       # split-K weight gradients meet in fp32 atomics: equal up to summation order
        return torch.equal(a, c) or float((a - c).abs().max() / c.abs().max()) < 1e-6

    out0, g0 = run(b)
    out1, g1 = run(b, prefetch_of=b)
    assert model._prefetched is None                                   # consumed
    assert torch.equal(out0, out1) and same_grads(g1, g0)
    out2, g2 = run(b, prefetch_of=other)                               # plan of another batch: not used for b
    assert torch.equal(out0, out2) and same_grads(g2, g0)
    out3, g3 = run(other)                                              # ... and still pending for `other`
    assert model._prefetched is None and out3.shape[0] == 3
    out4, g4 = run(b, prefetch_of=b, front="mol")
    assert torch.equal(out0, out4) and same_grads(g4, g0)
    for _ in range(20):                                                # back-to-back steps: buffers of the side stream's pool stay valid
        out5, g5 = run(b, prefetch_of=b)
        assert torch.equal(out0, out5) and same_grads(g5, g0)
    # host batch: copied to the device on the side stream too; the returned object is the one to call the model with
    host = synthetic_qm9_batch(8, seed=7).pin_memory()
    for _ in range(5):
        for p in model.parameters():
            p.grad = None
        nb = model.prefetch(host)
        assert nb.x.is_cuda and model._prefetched[0] is nb
        out6 = model(nb)
        (out6 - nb.y).abs().mean().backward()
        torch.cuda.synchronize()
        g6 = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        assert model._prefetched is None and torch.equal(out0, out6.detach()) and same_grads(g6, g0)
