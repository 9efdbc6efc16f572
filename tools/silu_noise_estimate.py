"""CPU estimate for the PAMNET_FAST_SILU experiment: the oracle (test infrastructure) in fp32 with a sigmoid perturbed by the
relative error of ex2.approx + rcp.approx, against fp64, on a QM9 dim=128 L=6 batch.  Result when written: output 8e-8,
worst parameter gradient 2.0e-6 (unchanged from plain fp32), no tensor near the 1e-5 ladder.  python tools/silu_noise_estimate.py"""
import sys, types, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pamnet_oracle as O
from pamnet_b200.data import synthetic_qm9_batch
torch.manual_seed(0)
cfg = types.SimpleNamespace(dataset="QM9", dim=128, n_layer=6, cutoff_l=5.0, cutoff_g=5.0, flow="source_to_target")
batch = synthetic_qm9_batch(8, seed=0)
sd = O.init_state_dict(cfg, seed=0)

def run(dtype, noisy):
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    orig = O.silu
    if noisy:
        g = torch.Generator().manual_seed(1)
        def silu(x):
            s = torch.sigmoid(x)
            # ex2.approx: ~2 ulp plus |x| * 2^-24 from the x*log2(e) product; rcp.approx: 1 ulp
            eps = (torch.rand(x.shape, generator=g, dtype=x.dtype) * 2 - 1) * (3.0 + x.abs()) * 2.0 ** -24
            return x * (s * (1 + eps.detach()))
        O.silu = silu
    try:
        out = O.forward(leaves, cfg, batch)
        (out - batch.y.to(dtype)).abs().mean().backward()
    finally:
        O.silu = orig
    return out.detach().double(), {k: v.grad.double() for k, v in leaves.items() if v.grad is not None}

o64, g64 = run(torch.float64, False)
o32, g32 = run(torch.float32, False)
on, gn = run(torch.float32, True)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
print("out  fp32 vs fp64 %.2e   noisy-silu fp32 vs fp64 %.2e" % (rel(o32, o64), rel(on, o64)))
w32 = max((rel(g32[k], g64[k]), k) for k in g64)
wn = max((rel(gn[k], g64[k]), k) for k in g64)
print("grad worst fp32 %.2e (%s)   noisy %.2e (%s)" % (w32[0], w32[1], wn[0], wn[1]))
bad = [(k, rel(gn[k], g64[k]), rel(g32[k], g64[k])) for k in g64 if rel(gn[k], g64[k]) > max(1e-5, 2 * rel(g32[k], g64[k]))]
print("tensors failing the ladder:", len(bad), bad[:5])
