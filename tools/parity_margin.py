"""GPU: margins of the parity ladder at BASELINE configs[1] size (test_full_size_vs_oracle[32-128-6]): worst tensors."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
from oracle import pamnet_oracle as O
from tests.helpers import ladder_ok, oracle_step
cfg = types.SimpleNamespace(dataset="QM9", dim=128, n_layer=6, cutoff_l=5.0, cutoff_g=5.0, flow="source_to_target")
sd = O.init_state_dict(cfg, seed=0)
b = synthetic_qm9_batch(32, seed=0)
model = PAMNet(Config(**vars(cfg))); model.load_state_dict(sd); model = model.cuda()
bd = b.to("cuda")
out = model(bd); (out - bd.y).abs().mean().backward(); torch.cuda.synchronize()
grads = {k: p.grad for k, p in model.named_parameters()}
cache = os.path.join(ROOT, "gpurun_out", "_oracle_c2.pt")
if os.path.exists(cache):
    o32, g32, o64, g64 = torch.load(cache)
else:
    o32, _, g32 = oracle_step(sd, cfg, b, dtype=torch.float32)
    o64, _, g64 = oracle_step(sd, cfg, b, dtype=torch.float64)
ok, e, er = ladder_ok(out, o32, o64)
rows = [("OUT", e, er, e / max(1e-5, 2 * er))]
for k, r64 in g64.items():
    if r64 is None: continue
    ok, e, er = ladder_ok(grads[k], g32[k], r64)
    rows.append((k, e, er, e / max(1e-5, 2 * er)))
rows.sort(key=lambda r: -r[3])
print("backend", os.environ.get("PAMNET_GEMM", "tc2"), "prod", os.environ.get("PAMNET_TC2_PROD", "4"), "chain", os.environ.get("PAMNET_CHAIN", "mma"))
for r in rows[:6]:
    print("   %-44s err %.2e  ref32 %.2e  frac-of-limit %.2f" % r)
