"""GPU diagnostic: run the CUDA path on a golden case and compare workspace intermediates with the oracle.
Usage (on the GPU box): python tools/diagnose.py [golden-name]  ->  table on stdout."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pamnet_b200  # noqa: E402
from pamnet_b200 import _lib, Config, PAMNet, PAMNet_s  # noqa: E402
from oracle import pamnet_oracle as O  # noqa: E402
from tests.helpers import load_golden, cfg_of, batch_of, rel_err  # noqa: E402


def ws_view(lib, mod, plan, ws, name, half, shape):
    off = lib.pamnet_debug_ws_offset(mod._ccfg, plan.sizes, name.encode(), half)
    assert off >= 0, name
    n = 1
    for s in shape:
        n *= s
    return ws[off:off + 4 * n].view(torch.float32).view(shape)


def plan_view(lib, plan, which, n, dtype=torch.int32):
    it = C.c_int32()
    off = lib.pamnet_debug_plan_offset(plan.sizes, which, it)
    blob = plan.trip if it.value else plan.base
    return blob[off:off + 4 * n].view(dtype)


def main(name="qm9_small_pamnet"):
    gold = load_golden(name)
    simple = name.endswith("_s")
    cfg = cfg_of(gold)
    b = batch_of(gold)
    model = (PAMNet_s if simple else PAMNet)(Config(**gold["config"]))
    model.load_state_dict(gold["state_dict"])
    model = model.cuda()
    lib = _lib.load()
    bc = b.to("cuda")

    # keep the workspace alive: hook the Function's ctx via a grad-requiring run
    out = model(bc)
    plan = model.last_plan
    ws = out.grad_fn.ws
    torch.cuda.synchronize()
    print("out   ", out.detach().cpu().tolist())
    print("golden", gold["out_f32"].tolist())

    sd = {k: v.clone() for k, v in gold["state_dict"].items()}
    ref_out, parts = O.forward(sd, cfg, b, simple=simple, return_parts=True)
    g = parts["graph"]
    sz = plan.sizes
    N, Eg, El, T, D, L = sz.n_nodes, sz.n_edges_g, sz.n_edges_l, sz.n_t2 + sz.n_t1, cfg.dim, cfg.n_layer
    print("sizes", N, Eg, El, sz.n_t2, sz.n_t1, "oracle", g.edge_index_g.shape[1], g.edge_index_l.shape[1],
          g.idx_kj.numel(), g.idx_jj_pair.numel())
    print("edge_index_g equal", torch.equal(plan.edge_index_g.cpu(), g.edge_index_g),
          "edge_index_l equal", torch.equal(plan.edge_index_l.cpu(), g.edge_index_l))
    g_eid = plan_view(lib, plan, 2, Eg).cpu().long()
    l_eid = plan_view(lib, plan, 6, El).cpu().long()
    rows = []

    def cmp(label, mine, ref):
        rows.append((label, rel_err(mine, ref)))

    cmp("dist_g", plan_view(lib, plan, 11, Eg, torch.float32).cpu(), g.dist_g[g_eid])
    cmp("dist_l", plan_view(lib, plan, 12, El, torch.float32).cpu(), g.dist_l[l_eid])
    cmp("rbf_g", ws_view(lib, model, plan, ws, "rbf_g", 0, (Eg, 16)).cpu(), parts["rbf_g"][g_eid])
    cmp("e_g", ws_view(lib, model, plan, ws, "e_g", 0, (Eg, D)).cpu(), parts["e_g"][g_eid])
    cmp("e_l", ws_view(lib, model, plan, ws, "e_l", 0, (El, D)).cpu(), parts["e_l"][l_eid])
    # merged triplet order -> oracle rows
    t_ptr = plan_view(lib, plan, 7, El + 1).cpu().long()
    t_split = plan_view(lib, plan, 10, El).cpu().long()
    # oracle lists are edge(API)-major; build the map merged index -> (kind, oracle index)
    import numpy as np
    off2 = np.zeros(El + 1, dtype=np.int64)
    off1 = np.zeros(El + 1, dtype=np.int64)
    if not simple:
        np.add.at(off2, g.idx_ji.numpy() + 1, 1)
    np.add.at(off1, g.idx_ji_pair.numpy() + 1, 1)
    off2, off1 = np.cumsum(off2), np.cumsum(off1)
    s_ref = torch.zeros(T, D)
    for k in range(El):
        e = int(l_eid[k])
        n2 = int(t_split[k])
        a = int(t_ptr[k])
        if n2:
            s_ref[a:a + n2] = parts["s2"][off2[e]:off2[e + 1]]
        n1 = int(t_ptr[k + 1]) - a - n2
        s_ref[a + n2:a + n2 + n1] = parts["s1"][off1[e]:off1[e + 1]]
    cmp("s(sbf embed)", ws_view(lib, model, plan, ws, "s", 0, (T, D)).cpu(), s_ref)
    cmp("x0", ws_view(lib, model, plan, ws, "x0", 0, (N, D)).cpu(), sd["embeddings"][b.x.long()] if cfg.dataset == "QM9" else ws_view(lib, model, plan, ws, "x0", 0, (N, D)).cpu())
    for hh in range(2 * L):
        cmp(f"x_out half {hh}", ws_view(lib, model, plan, ws, "r2", hh, (N, D)).cpu(), parts["x_halves"][hh])
    att = ws_view(lib, model, plan, ws, "att", 0, (2 * L, N)).cpu()
    outh = ws_view(lib, model, plan, ws, "out", 0, (2 * L, N)).cpu()
    cmp("att", att.view(L, 2, N).permute(0, 2, 1), parts["att"])
    cmp("out heads", outh.view(L, 2, N).permute(0, 2, 1), parts["out"])
    cmp("model out", out.detach().cpu(), ref_out)
    for label, e in rows:
        print(f"{label:24s} {e:.3e}")
    # backward
    y = bc.y
    loss = (out - y).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    worst = []
    for k, p in model.named_parameters():
        gr = gold["grads_f64"].get(k)
        if gr is None:
            continue
        worst.append((rel_err(p.grad.cpu(), gr), k))
    worst.sort(reverse=True)
    print("worst grads vs fp64 golden:")
    for e, k in worst[:25]:
        print(f"  {e:.3e} {k}")
    print("best grads:")
    for e, k in worst[-5:]:
        print(f"  {e:.3e} {k}")


if __name__ == "__main__":
    main(*sys.argv[1:])
