"""GPU: operator-surface kernels (through the C ABI) against the oracle's graph ops / torch fp32."""
import pytest
import torch

from tests.helpers import load_golden, batch_of, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from pamnet_b200 import ops
    return ops


def _qm9(n=16, seed=3):
    from pamnet_b200.data import synthetic_qm9_batch
    return synthetic_qm9_batch(n, seed=seed)


def test_native_library_is_loaded(ops):
    from pamnet_b200 import _lib
    lib = _lib.load()
    assert lib.pamnet_abi_version() == 1
    with open("/proc/self/maps") as f:
        assert "libpamnet_sm100.so" in f.read()


@pytest.mark.parametrize("r,max_nb", [(5.0, 1000), (2.0, 1000), (5.0, 3), (0.5, 32)])
def test_radius_bit_exact(ops, r, max_nb):
    from oracle import graph_ops as G
    b = _qm9()
    row, col = G.radius_pairs(b.pos, b.pos, r, b.batch, b.batch, max_nb)
    mine = ops.radius(b.pos.cuda(), b.pos.cuda(), r, b.batch.cuda(), b.batch.cuda(), max_nb)
    assert torch.equal(mine[0].cpu(), row) and torch.equal(mine[1].cpu(), col)
    eg = ops.radius_graph(b.pos.cuda(), b.batch.cuda(), r, max_nb, loop=False)
    assert torch.equal(eg.cpu(), G.drop_self_loops(torch.stack([row, col])))


def test_radius_golden(ops):
    gold = load_golden("qm9_small_pamnet")
    b = batch_of(gold)
    row, col = ops.radius(b.pos.cuda(), b.pos.cuda(), 5.0, b.batch.cuda(), b.batch.cuda(), 1000)
    assert torch.equal(row.cpu(), gold["graph"]["radius_row"]) and torch.equal(col.cpu(), gold["graph"]["radius_col"])


def test_radius_edge_cases(ops):
    dev = "cuda"
    pos = torch.tensor([[0.0, 0, 0]], device=dev)
    row, col = ops.radius(pos, pos, 1.0, torch.zeros(1, dtype=torch.long, device=dev), None, 10)
    assert row.tolist() == [0] and col.tolist() == [0]                 # single atom: only itself
    assert ops.radius_graph(pos, torch.zeros(1, dtype=torch.long, device=dev), 1.0, 10).shape == (2, 0)
    # two graphs that overlap in space must not connect
    pos = torch.tensor([[0.0, 0, 0], [0.1, 0, 0], [0.05, 0, 0]], device=dev)
    batch = torch.tensor([0, 0, 1], device=dev)
    ei = ops.radius_graph(pos, batch, 1.0, 10)
    assert ei.cpu().tolist() == [[0, 1], [1, 0]]
    # exact tie on the cutoff is kept (d2 <= r2)
    pos = torch.tensor([[0.0, 0, 0], [3.0, 4.0, 0]], device=dev)
    assert ops.radius_graph(pos, torch.zeros(2, dtype=torch.long, device=dev), 5.0, 10).shape[1] == 2


@pytest.mark.parametrize("k", [4, 50])
def test_knn_bit_exact(ops, k):
    from oracle import graph_ops as G
    from pamnet_b200.data import synthetic_rna_batch
    b = synthetic_rna_batch(2, seed=1, min_atoms=40, max_atoms=90)
    pos = b.x[:, :3].contiguous()
    row, col = G.knn_pairs(pos, pos, k, b.batch, b.batch)
    mine = ops.knn(pos.cuda(), pos.cuda(), k, b.batch.cuda(), b.batch.cuda())
    assert torch.equal(mine[0].cpu(), row) and torch.equal(mine[1].cpu(), col)


def test_knn_small_graph_and_ties(ops):
    pos = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [-1.0, 0, 0], [0, 2.0, 0]], device="cuda")
    batch = torch.zeros(4, dtype=torch.long, device="cuda")
    row, col = ops.knn(pos, pos, 50, batch, batch)          # fewer than k points: all of them, by distance
    assert row.cpu().tolist() == [0] * 4 + [1] * 4 + [2] * 4 + [3] * 4
    assert col.cpu().tolist()[:4] == [0, 1, 2, 3]           # tie between 1 and 2 -> lower index first


def test_remove_self_loops_and_filter(ops):
    from oracle import graph_ops as G
    b = _qm9(6)
    ei = b.edge_index.clone()
    ei[:, 3] = ei[0, 3]
    ei[:, 10] = ei[1, 10]
    out, _ = ops.remove_self_loops(ei.cuda())
    assert torch.equal(out.cpu(), G.drop_self_loops(ei))
    kept = ops.filter_edges(b.edge_index.cuda(), b.pos.cuda(), 1.3)
    ref = b.edge_index[:, G.edge_lengths(b.edge_index, b.pos) <= 1.3]
    assert torch.equal(kept.cpu(), ref)
    assert ops.filter_edges(torch.zeros((2, 0), dtype=torch.long, device="cuda")).shape == (2, 0)


@pytest.mark.parametrize("seed", [0, 7])
def test_triplet_indices_bit_exact(ops, seed):
    from oracle import graph_ops as G
    b = _qm9(8, seed)
    ref = G.triplet_indices(b.edge_index, b.pos.shape[0])
    mine = ops.triplet_indices(b.edge_index.cuda(), b.pos.shape[0])
    for name, r, m in zip(ops.TRIPLET_NAMES, ref, mine):
        assert torch.equal(m.cpu(), r), name
    # shuffled (API order is arbitrary) and asymmetric edge lists
    perm = torch.randperm(b.edge_index.shape[1], generator=torch.Generator().manual_seed(seed))
    ei = b.edge_index[:, perm][:, : b.edge_index.shape[1] * 2 // 3]
    for r, m in zip(G.triplet_indices(ei, b.pos.shape[0]), ops.triplet_indices(ei.cuda(), b.pos.shape[0])):
        assert torch.equal(m.cpu(), r)


def test_triplet_indices_golden_and_empty(ops):
    gold = load_golden("qm9_small_pamnet")
    gv = gold["graph"]
    mine = ops.triplet_indices(gv["edge_index_l"].cuda(), batch_of(gold).pos.shape[0])
    for name, m in zip(ops.TRIPLET_NAMES, mine):
        assert torch.equal(m.cpu(), gv[name]), name
    empty = ops.triplet_indices(torch.zeros((2, 0), dtype=torch.long, device="cuda"), 5)
    assert all(t.numel() == 0 for t in empty)


def test_scatter_add(ops):
    g = torch.Generator().manual_seed(0)
    src = torch.randn(1000, 48, generator=g)
    idx = torch.randint(0, 77, (1000,), generator=g)
    ref = torch.zeros(80, 48).index_add_(0, idx, src)
    out = ops.scatter(src.cuda(), idx.cuda(), dim=0, dim_size=80)
    assert rel_err(out, ref) < 1e-6
    assert ops.scatter(src[:0].cuda(), idx[:0].cuda(), dim_size=3).abs().sum() == 0


def test_bessel_and_spherical_basis(ops):
    from oracle import pamnet_oracle as O
    g = torch.Generator().manual_seed(1)
    dist = torch.rand(500, generator=g) * 4.0 + 0.9
    freq = torch.arange(1, 17, dtype=torch.float32) * 3.14159265 + torch.randn(16, generator=g) * 0.01
    ref = O.bessel_rbf(dist.double(), freq.double(), 5.0)
    assert rel_err(ops.bessel_rbf(dist.cuda(), freq.cuda(), 5.0), ref) < 2e-6
    angle = torch.rand(2000, generator=g) * 3.14159
    gather = torch.randint(0, 500, (2000,), generator=g)
    ref = O.spherical_basis(dist.double(), angle.double(), gather, 5.0)
    out = ops.spherical_basis(dist.cuda(), angle.cuda(), gather.cuda(), 5.0)
    assert rel_err(out, ref) < 2e-6          # the fp32 reference itself is only good to ~7e-5 here


@pytest.mark.parametrize("n_in,n_out,rows", [(16, 128, 777), (42, 16, 100), (128, 128, 1), (384, 128, 333), (18, 32, 65),
                                              (16, 16, 9001), (32, 16, 4100)])
def test_linear(ops, n_in, n_out, rows):
    g = torch.Generator().manual_seed(2)
    x, w, b = torch.randn(rows, n_in, generator=g), torch.randn(n_out, n_in, generator=g) / n_in ** 0.5, torch.randn(n_out, generator=g)
    ref = x.double() @ w.double().T + b.double()
    assert rel_err(ops.linear(x.cuda(), w.cuda(), b.cuda()), ref) < 4e-6
    assert rel_err(ops.linear(x.cuda(), w.cuda(), b.cuda(), silu=True), ref * torch.sigmoid(ref)) < 4e-6
    assert rel_err(ops.linear(x.cuda(), w.cuda()), x.double() @ w.double().T) < 4e-6


@pytest.mark.parametrize("m,n,k,ksplit", [(128, 128, 599, 1), (128, 16, 5000, 4), (1, 128, 300, 1), (70, 88, 1000, 3),
                                            (300, 128, 128, 1), (1000, 256, 64, 1), (128, 384, 2000, 8), (130, 100, 16, 1),
                                            (257, 129, 40, 2),
                                            # skinny weight gradients (gemm_small.cu column kernel: register fp64 accumulators)
                                            (16, 16, 100003, 1), (32, 16, 7777, 1), (16, 88, 5000, 1), (32, 32, 3001, 1),
                                            (32, 96, 1000, 1), (3, 5, 17, 1), (9, 13, 40000, 1),
                                            # dim-16 rows kernels on the tensor cores (N = 16, K = 16 / 32, many rows)
                                            (5000, 16, 16, 1), (70001, 16, 32, 1)])
def test_gemm_modes(ops, m, n, k, ksplit):
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g)
    # fp32 running sums of k unit-variance terms: rounding error grows ~ eps * k / sqrt(2) in absolute terms
    # (tensor-core path: 3xTF32 keeps ~2^-21 relative error per product -> ~5e-7 * sqrt(k) more)
    tol = 2e-7 * k + 1e-5 * k ** 0.5

    def err(x, ref):
        return float((x.double().cpu() - ref).abs().max())
    assert err(ops.gemm(0, a.cuda(), b.cuda(), m, n, k), a.double() @ b.double().T) < tol
    bt = b.T.contiguous()
    assert err(ops.gemm(1, a.cuda(), bt.cuda(), m, n, k), a.double() @ bt.double()) < tol
    at = a.T.contiguous()
    c, db = ops.gemm(2, at.cuda(), bt.cuda(), m, n, k, ksplit=ksplit, want_dbias=True)
    assert err(c, at.double().T @ bt.double()) < tol
    assert err(db, at.double().sum(0)) < tol


def test_fused_losses_match_torch(ops):
    """ops.l1_loss / ops.mse_loss (pamnet_loss: value and gradient in one launch) against F.l1_loss / F.mse_loss."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    for n in (1, 32, 1000):
        y = torch.randn(n, device="cuda")
        for fused, ref in ((ops.l1_loss, F.l1_loss), (ops.mse_loss, F.mse_loss)):
            a = torch.randn(n, device="cuda", requires_grad=True)
            b = a.detach().clone().requires_grad_(True)
            la, lb = fused(a, y), ref(b, y)
            (3.0 * la).backward()
            (3.0 * lb).backward()
            assert abs(float(la) - float(lb)) <= 1e-6 * max(1.0, abs(float(lb)))
            assert torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-9)
    with pytest.raises(ValueError):
        ops.l1_loss(torch.zeros(3, device="cuda"), torch.zeros(4, device="cuda"))


@pytest.mark.parametrize("n_graphs,n_atoms,r,max_nb", [(3, 700, 6.0, 1000), (1, 2500, 2.6, 1000), (5, 40, 5.0, 1000),
                                                       (2, 900, 6.0, 20), (4, 300, 50.0, 64)])
def test_radius_grid_equals_brute_force_and_oracle(ops, n_graphs, n_atoms, r, max_nb):
    """Cell-list radius search (csrc/graph_grid.cu) against the per-graph scan and the oracle: identical edge lists, also
    with a binding max_num_neighbors cut, coincident points and a cutoff larger than the structure."""
    from oracle import graph_ops as G
    g = torch.Generator().manual_seed(n_atoms)
    pos = torch.rand(n_graphs * n_atoms, 3, generator=g) * (n_atoms ** (1 / 3.0)) * 2.2        # ~0.1 atoms / A^3
    pos[5] = pos[4]                                                                              # coincident atoms
    pos = torch.round(pos * 1000) / 1000
    batch = torch.arange(n_graphs).repeat_interleave(n_atoms)
    for loop in (False, True):
        brute = ops.radius_graph(pos.cuda(), batch.cuda(), r, max_nb, loop=loop, method="brute")
        grid = ops.radius_graph(pos.cuda(), batch.cuda(), r, max_nb, loop=loop, method="grid")
        assert torch.equal(brute, grid)
    row, col = G.radius_pairs(pos, pos, r, batch, batch, max_nb)
    ref = torch.stack([row, col])
    assert torch.equal(ops.radius_graph(pos.cuda(), batch.cuda(), r, max_nb, loop=True, method="grid").cpu(), ref)
