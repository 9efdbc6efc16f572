"""GPU: BASELINE.json configs[3] -- rna dim=16 n_layer=1, 8 graphs of ~2k atoms (kNN-50 global graph, target_to_source):
ms per forward + L1 loss + backward and the kernel-class breakdown.  python tools/rna_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import pamnet_b200
from pamnet_b200 import Config, PAMNet, _lib
from pamnet_b200.data import Batch, synthetic_rna_batch

torch.manual_seed(0)
model = PAMNet(Config("rna_native", 16, 1, 2.6, 20.0, "target_to_source")).cuda()
base = synthetic_rna_batch(8, seed=0, min_atoms=300, max_atoms=500)
xs, bs = [], []
for g in range(8):                      # ~2k-atom graphs: five translated copies of a 300-500 atom chain
    xg = base.x[base.batch == g]
    for c in range(5):
        xs.append(xg + torch.tensor([37.0 * c, 11.0 * c, 0.0, 0.0]))
        bs.append(torch.full((xg.shape[0],), g, dtype=torch.long))
b = Batch(x=torch.cat(xs), batch=torch.cat(bs), y=base.y).to("cuda")
params = list(model.parameters())

def step():
    for p in params:
        p.grad = None
    F.l1_loss(model(b), b.y).backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
sz = model.last_plan.sizes
print(f"rna bs=8: N={sz.n_nodes} E_g={sz.n_edges_g} E_l={sz.n_edges_l} T2={sz.n_t2} T1={sz.n_t1}: {ms:.3f} ms/step = {8 / ms * 1e3:.0f} graphs/s")
_lib.profile_begin()
for _ in range(5):
    step()
prof = _lib.profile_end()
tot = sum(v[0] for v in prof.values())
print({k: (round(v[0] / 5, 3), v[1] // 5) for k, v in prof.items() if v[1]})
