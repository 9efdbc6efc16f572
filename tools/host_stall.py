"""GPU: where does the host lose time in the rare slow steps of the prefetching loop?  Host timestamps of both threads."""
import os, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
if os.environ.get("SWITCH"):
    sys.setswitchinterval(float(os.environ["SWITCH"]))
torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")
wlog = []
orig = model.prefetch
def timed_prefetch(*a, **k):
    t0 = time.perf_counter()
    r = orig(*a, **k)
    wlog.append((t0, time.perf_counter()))
    return r
model.prefetch = timed_prefetch
fut = []
rows = []
N = int(os.environ.get("STEPS", "2000"))
gc.collect(); gc.disable()
for i in range(N + 20):
    t0 = time.perf_counter()
    if fut:
        fut.pop(0).result()
    t1 = time.perf_counter()
    model.zero_grad()
    out = model(b)
    t2 = time.perf_counter()
    fut.append(model.prefetch_async(b, wait_current=False))
    t3 = time.perf_counter()
    loss = pamnet_b200.ops.l1_loss(out, b.y)
    t4 = time.perf_counter()
    loss.backward()
    t5 = time.perf_counter()
    if i % 50 == 49:
        torch.cuda.synchronize()
    rows.append((t0, t1, t2, t3, t4, t5, time.perf_counter()))
torch.cuda.synchronize()
rows = rows[20:]
tot = sorted(r[6] - r[0] for r in rows)
print("host step: median %.3f ms  p99 %.3f  max %.3f" % (1e3 * tot[len(tot) // 2], 1e3 * tot[int(len(tot) * .99)], 1e3 * tot[-1]))
names = ["result()", "zero+forward", "submit", "loss", "backward", "sync"]
for i, r in enumerate(rows):
    if r[6] - r[0] > 4e-3:
        parts = {n: round(1e3 * (r[j + 1] - r[j]), 2) for j, n in enumerate(names)}
        w = [(round(1e3 * (a - r[0]), 2), round(1e3 * (e - a), 2)) for a, e in wlog if a < r[6] and e > r[0]]
        print("step", i, "total %.2f ms" % (1e3 * (r[6] - r[0])), parts, "worker (start rel, dur):", w)
wd = sorted(e - a for a, e in wlog)
print("worker prefetch: median %.3f ms  p99 %.3f  max %.3f" % (1e3 * wd[len(wd) // 2], 1e3 * wd[int(len(wd) * .99)], 1e3 * wd[-1]))
