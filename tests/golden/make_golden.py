"""Regenerate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/models.py verbatim through oracle/ref_shim.py (stand-ins for the
four absent third-party packages), runs it on seeded synthetic inputs and on the shipped RNA
checkpoint x native structures, and stores inputs, weights, outputs (fp32 and fp64 rungs),
parameter gradients and the integer graph/index vectors.  The fixtures travel to the GPU box;
/root/reference does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import pamnet_b200  # noqa: E402
from pamnet_b200.data import Batch, synthetic_qm9_batch  # noqa: E402
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference(model, batch, loss="l1"):
    """fp32 forward+backward and fp64 forward+backward of the verbatim reference."""
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = model.to(dt)
        m.zero_grad()
        fields = dict(batch.__dict__)
        for k in ("pos",):
            if k in fields:
                fields[k] = fields[k].to(dt)
        if fields["x"].dim() == 2:
            fields["x"] = fields["x"].to(dt)
        b = Batch(**fields)
        y = m(b)
        target = batch.y.to(dt)
        l = (y - target).abs().mean() if loss == "l1" else ((y - target) ** 2).mean()
        l.backward()
        out["out_" + tag] = y.detach().clone()
        out["loss_" + tag] = l.detach().clone()
        out["grads_" + tag] = {k: (p.grad.detach().clone() if p.grad is not None else None)
                               for k, p in m.named_parameters()}
    model.to(torch.float32)
    return out


def graph_vectors(ref, model, batch):
    """edge lists and the ten index vectors exactly as the reference derives them."""
    from torch_geometric.nn import radius
    pos = batch.pos
    row, col = radius(pos, pos, model.cutoff_g, batch.batch, batch.batch, max_num_neighbors=1000)
    eg, dist_g = model.get_edge_info(torch.stack([row, col]), pos)
    el, dist_l = model.get_edge_info(batch.edge_index, pos)
    names = ["idx_i", "idx_j", "idx_k", "idx_kj", "idx_ji", "idx_i_pair", "idx_j1_pair", "idx_j2_pair",
             "idx_jj_pair", "idx_ji_pair"]
    idx = dict(zip(names, model.indices(el, num_nodes=pos.shape[0])))
    return dict(radius_row=row, radius_col=col, edge_index_g=eg, edge_index_l=el, dist_g=dist_g,
                dist_l=dist_l, **idx)


def make_qm9(ref):
    for name, cls, seed in (("qm9_small_pamnet", ref.PAMNet, 11), ("qm9_small_pamnet_s", ref.PAMNet_s, 12)):
        torch.manual_seed(seed)
        cfg = ref.Config("QM9", 32, 2, 5.0, 5.0)
        model = cls(cfg)
        batch = synthetic_qm9_batch(4, seed=seed)
        res = run_reference(model, batch)
        res["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
        res["batch"] = dict(batch.__dict__)
        res["config"] = dict(dataset="QM9", dim=32, n_layer=2, cutoff_l=5.0, cutoff_g=5.0,
                             flow="source_to_target")
        if cls is ref.PAMNet:
            res["graph"] = graph_vectors(ref, model, batch)
        torch.save(res, os.path.join(HERE, name + ".pt"))
        print(name, res["out_f32"])


def read_tu(root, name):
    """TU text format (reference datasets/tu_dataset.py:90-122): node_attributes + node_labels."""
    raw = os.path.join(root, name, "raw")
    attrs = np.loadtxt(os.path.join(raw, name + "_node_attributes.txt"), delimiter=",", dtype=np.float64)
    labels = np.loadtxt(os.path.join(raw, name + "_node_labels.txt"), dtype=np.int64)
    gid = np.loadtxt(os.path.join(raw, name + "_graph_indicator.txt"), dtype=np.int64) - 1
    y = np.loadtxt(os.path.join(raw, name + "_graph_labels.txt"), delimiter=",", dtype=np.float64)
    labels = labels - labels.min()
    x = np.concatenate([attrs, labels[:, None].astype(np.float64)], 1).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(gid), torch.from_numpy(np.atleast_1d(y).astype(np.float32))


def make_rna(ref, ref_root):
    x, gid, y = read_tu(os.path.join(ref_root, "data", "RNA-Puzzles"), "rna_native")
    cfg = ref.Config("rna_native", 16, 1, 2.6, 20.0, "target_to_source")
    model = ref.PAMNet(cfg)
    sd = torch.load(os.path.join(ref_root, "save", "pamnet_rna.pt"), map_location="cpu")
    print(model.load_state_dict(sd))
    model.eval()
    n_graphs = int(gid.max()) + 1
    scores32, scores64, sizes = [], [], []
    for g in range(n_graphs):
        sel = gid == g
        b = Batch(x=x[sel], batch=torch.zeros(int(sel.sum()), dtype=torch.long), y=y[g:g + 1])
        scores32.append(model(b).detach())
        m64 = model.double()
        scores64.append(m64(Batch(x=x[sel].double(), batch=b.batch, y=b.y)).detach())
        model.float()
        sizes.append(int(sel.sum()))
        print(g, sizes[-1], float(scores32[-1]), float(scores64[-1]))
    order = np.argsort(sizes)[:2]                       # ship the two smallest structures as inputs
    inputs = {int(g): x[gid == int(g)].clone() for g in order}
    # one training-style fwd+bwd on the two shipped graphs batched together
    xb = torch.cat([inputs[int(g)] for g in order])
    bb = torch.cat([torch.full((inputs[int(g)].shape[0],), i, dtype=torch.long) for i, g in enumerate(order)])
    batch = Batch(x=xb, batch=bb, y=torch.tensor([1.5, 4.0]))
    res = run_reference(model, batch, loss="l1")
    res.update(state_dict={k: v.clone() for k, v in sd.items()}, scores_f32=torch.cat(scores32),
               scores_f64=torch.cat(scores64), sizes=sizes, shipped=[int(g) for g in order], inputs=inputs,
               batch=dict(batch.__dict__),
               config=dict(dataset="rna_native", dim=16, n_layer=1, cutoff_l=2.6, cutoff_g=20.0,
                           flow="target_to_source"))
    torch.save(res, os.path.join(HERE, "rna_native.pt"))


def make_rna_c4(ref, ref_root):
    """BASELINE.json configs[3]: the FIRST 8 graphs of the rna_native fixture (841-3 771 atoms,
    15 816 in total) batched together, shipped checkpoint, one training-style forward + L1 loss + backward."""
    x, gid, y = read_tu(os.path.join(ref_root, "data", "RNA-Puzzles"), "rna_native")
    cfg = ref.Config("rna_native", 16, 1, 2.6, 20.0, "target_to_source")
    model = ref.PAMNet(cfg)
    sd = torch.load(os.path.join(ref_root, "save", "pamnet_rna.pt"), map_location="cpu")
    model.load_state_dict(sd)
    sel = gid < 8
    batch = Batch(x=x[sel].clone(), batch=gid[sel].clone(), y=torch.linspace(0.5, 11.0, 8))
    res = run_reference(model, batch, loss="l1")
    sizes = [int((gid == g).sum()) for g in range(8)]
    print("C4 sizes", sizes, "out", res["out_f32"])
    res.update(state_dict={k: v.clone() for k, v in sd.items()}, sizes=sizes,
               x=batch.x, y=batch.y,            # batch vector = repeat_interleave(arange(8), sizes)
               config=dict(dataset="rna_native", dim=16, n_layer=1, cutoff_l=2.6, cutoff_g=20.0,
                           flow="target_to_source"))
    torch.save(res, os.path.join(HERE, "rna_c4.pt"))


def make_pdbbind(ref):
    torch.manual_seed(21)
    rng = np.random.default_rng(21)
    cfg = ref.Config("PDBbind", 16, 1, 2.0, 6.0)
    model = ref.PAMNet(cfg)
    xs, bs = [], []
    for g in range(2):
        mol = synthetic_qm9_batch(3, seed=100 + g)      # three fragments -> one "complex"
        pos = mol.pos.numpy().astype(np.float64)
        pos[mol.batch.numpy() == 1] += 3.0
        pos[mol.batch.numpy() == 2] += np.array([45.0, 0, 0])   # x > 40 marks the sign flip (models.py:125)
        feats = rng.normal(size=(pos.shape[0], 18))
        xs.append(np.concatenate([pos, feats], 1).astype(np.float32))
        bs.append(np.full(pos.shape[0], g, dtype=np.int64))
    batch = Batch(x=torch.from_numpy(np.concatenate(xs)), batch=torch.from_numpy(np.concatenate(bs)),
                  y=torch.tensor([5.0, 7.0]))
    res = run_reference(model, batch, loss="mse")
    res["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
    res["batch"] = dict(batch.__dict__)
    res["config"] = dict(dataset="PDBbind", dim=16, n_layer=1, cutoff_l=2.0, cutoff_g=6.0,
                         flow="source_to_target")
    torch.save(res, os.path.join(HERE, "pdbbind_small.pt"))
    print("pdbbind", res["out_f32"])


if __name__ == "__main__":
    ref_root = os.environ.get("PAMNET_REFERENCE_ROOT", ref_shim.REF_ROOT_DEFAULT)
    ref = ref_shim.load_reference(ref_root)
    make_qm9(ref)
    make_pdbbind(ref)
    make_rna(ref, ref_root)
    make_rna_c4(ref, ref_root)
