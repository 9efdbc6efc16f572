"""GPU: what does one NVML clock / throttle-reason query cost while a training loop runs (host time of the call itself)?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml as nv
import pamnet_b200
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")
def step():
    model.zero_grad(); out = model(b); pamnet_b200.ops.l1_loss(out, b.y).backward()
for _ in range(20): step()
torch.cuda.synchronize()
calls = {"clock": lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
         "reasons": lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h),
         "maxclock": lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)}
for name, fn in calls.items():
    ts = []
    for i in range(300):
        step()
        if i % 10 == 5:
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    ts.sort()
    print(f"{name}: main-thread call median {1e3*ts[len(ts)//2]:.3f} ms  max {1e3*ts[-1]:.3f} ms  (n={len(ts)})")
# step-time effect of a background sampler calling one query every 50 ms
for name, fn in calls.items():
    stop = threading.Event()
    def loop():
        while not stop.is_set():
            fn(); stop.wait(0.05)
    th = threading.Thread(target=loop, daemon=True); th.start()
    ts = []
    for i in range(400):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
        if i % 50 == 49: torch.cuda.synchronize()
    stop.set(); th.join()
    torch.cuda.synchronize()
    ts.sort()
    print(f"background {name} @20 Hz: host step median {1e3*ts[len(ts)//2]:.3f} ms  p99 {1e3*ts[int(len(ts)*.99)]:.3f}  max {1e3*ts[-1]:.3f}")
