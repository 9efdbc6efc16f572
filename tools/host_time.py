"""GPU: host enqueue time of one training step (no synchronisation inside the loop) vs the device time.
    python tools/host_time.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet, ops
from pamnet_b200.data import synthetic_qm9_batch
torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")
params = list(model.parameters())
def step(parts):
    t0 = time.perf_counter()
    model.zero_grad()
    t1 = time.perf_counter()
    out = model(b)
    t2 = time.perf_counter()
    loss = ops.l1_loss(out, b.y)
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    model.prefetch(b, wait_current=False)
    t5 = time.perf_counter()
    for i, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): parts[i] += d
for _ in range(10): step([0] * 5)
torch.cuda.synchronize()
import gc; gc.disable()
N = 200
parts = [0.0] * 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N): step(parts)
t_host = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("host enqueue per step %.3f ms; device span per step %.3f ms" % (1e3 * t_host / N, e0.elapsed_time(e1) / N))
print("  zero_grad %.3f  forward %.3f  loss %.3f  backward %.3f  prefetch %.3f (ms)" % tuple(1e3 * p / N for p in parts))
