"""Device-side collation (csrc/collate.cuh, DeviceDataset) on the CPU: the kernel body compiled for the host against the
collate a PyG DataLoader performs (concatenate, offset edge_index, graph id per atom), and the host-side table logic."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pamnet_b200.data import DeviceDataset, molecules_of, synthetic_qm9_batch  # noqa: E402


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    src = os.path.join(ROOT, "tests", "host_emul", "collate_host.cpp")
    out = str(tmp_path_factory.mktemp("collate") / "collate_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def _reference_collate(mols, ids):
    xs, ps, es, bs, off = [], [], [], [], 0
    for g, i in enumerate(ids):
        m = mols[i]
        xs.append(m.x); ps.append(m.pos); es.append(m.edge_index + off); bs.append(np.full(m.x.shape[0], g))
        off += m.x.shape[0]
    return (np.concatenate(xs), np.concatenate(ps), np.concatenate(es, axis=1), np.concatenate(bs),
            np.array([mols[i].y for i in ids], dtype=np.float32))


def test_molecules_of_round_trip():
    b = synthetic_qm9_batch(6, seed=3)
    mols = molecules_of(b)
    x, pos, ei, bv, y = _reference_collate(mols, list(range(6)))
    assert np.array_equal(x, b.x.numpy()) and np.array_equal(pos, b.pos.numpy()) and np.array_equal(bv, b.batch.numpy())
    assert np.array_equal(ei, b.edge_index.numpy()) and np.allclose(y, b.y.numpy())


def test_collate_kernel_body_matches_loader_collate(host_lib):
    mols = molecules_of(synthetic_qm9_batch(12, seed=1))
    mols[4].edge_index = np.zeros((2, 0), dtype=np.int64)            # a molecule without bonds
    ds = DeviceDataset.__new__(DeviceDataset)                          # host-side fields only (no GPU here)
    ds.n_atoms = np.array([m.x.shape[0] for m in mols], dtype=np.int64)
    ds.n_bonds = np.array([m.edge_index.shape[1] for m in mols], dtype=np.int64)
    node_ptr = np.concatenate([[0], np.cumsum(ds.n_atoms)]).astype(np.int64)
    edge_ptr = np.concatenate([[0], np.cumsum(ds.n_bonds)]).astype(np.int64)
    x_all = np.concatenate([m.x for m in mols]).astype(np.float32)
    pos_all = np.ascontiguousarray(np.concatenate([m.pos for m in mols]).astype(np.float32))
    ei_all = np.ascontiguousarray(np.concatenate([m.edge_index for m in mols], axis=1).astype(np.int64))
    y_all = np.array([m.y for m in mols], dtype=np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for ids in ([0, 1, 2], [11, 4, 4, 7, 0], [5]):
        tab, n, e = ds.table(ids)
        tab = np.ascontiguousarray(tab)
        x, pos = np.full(n, np.nan, np.float32), np.full((n, 3), np.nan, np.float32)
        ei, bv, y = np.full((2, e), -1, np.int64), np.full(n, -1, np.int64), np.full(len(ids), np.nan, np.float32)
        host_lib.collate_host(p(tab), C.c_int64(len(ids)), p(node_ptr), p(edge_ptr), p(x_all), p(pos_all), p(ei_all),
                              C.c_int64(ei_all.shape[1]), p(y_all), C.c_int64(e), p(x), p(pos), p(ei), p(bv), p(y))
        rx, rp, re, rb, ry = _reference_collate(mols, ids)
        assert np.array_equal(x, rx) and np.array_equal(pos, rp) and np.array_equal(ei, re)
        assert np.array_equal(bv, rb) and np.array_equal(y, ry)
    with pytest.raises(IndexError):
        ds.table([12])


def test_device_dataset_needs_a_gpu():
    from pamnet_b200 import PamnetError
    with pytest.raises(PamnetError):
        DeviceDataset(molecules_of(synthetic_qm9_batch(2, seed=0)), "cpu")


@pytest.mark.gpu
def test_device_dataset_batch_equals_loader_collate_gpu():
    from pamnet_b200 import Config, PAMNet
    ref = synthetic_qm9_batch(16, seed=2)
    mols = molecules_of(ref)
    ds = DeviceDataset(mols, "cuda")
    b = ds.batch(list(range(16)))
    for k in ("x", "pos", "edge_index", "batch", "y"):
        assert torch.equal(getattr(b, k).cpu(), getattr(ref, k)), k
    ids = [3, 3, 9, 0]
    b = ds.batch(ids)
    rx, rp, re, rb, ry = _reference_collate(mols, ids)
    assert np.array_equal(b.x.cpu().numpy(), rx) and np.array_equal(b.edge_index.cpu().numpy(), re)
    assert np.array_equal(b.batch.cpu().numpy(), rb) and b.num_graphs == 4
    torch.manual_seed(0)
    model = PAMNet(Config("QM9", 32, 1, 5.0, 5.0)).cuda()
    with torch.no_grad():
        assert torch.equal(model(ds.batch(list(range(16)))), model(ref.to("cuda")))
