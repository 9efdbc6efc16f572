"""Where does the host time of a step go?  (GPU box)  python tools/pyprofile.py"""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch

torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")

params = list(model.parameters())

def step():
    for p in params:
        p.grad = None
    out = model(b)
    loss = (out - b.y).abs().mean()
    loss.backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time/step {1e3*(t1-t0)/50:.3f} ms; incl. final drain {1e3*(t2-t0)/50:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
