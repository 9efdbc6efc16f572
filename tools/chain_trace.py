"""GPU: clock64 timeline of the node chain kernel (CTA 0) for the LAST chain launch of a training step.
    PAMNET_TC_TRACE=1 python physics-aware-multiplex-gnn_b200/build.py --force && python tools/chain_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet, _lib
from pamnet_b200.data import synthetic_qm9_batch

lib = _lib.load()
torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")
for which in ("forward", "backward"):
    for _ in range(3):
        for p in model.parameters():
            p.grad = None
        out = model(b)
        if which == "backward":
            (out - b.y).abs().mean().backward()
        torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 256)()
    _lib.check(lib.pamnet_debug_chain_trace(buf, 256), "trace")
    t = list(buf)
    print(f"last chain launch of {which}: per stage deltas: [scan, issue_w, bias/addg, prologue] weights barrier loop exchange [reduce, math+stores, slot] end")
    for si in range(16):
        r = t[si * 16:(si + 1) * 16]
        if r[0] == 0 or r[1] == 0:
            continue
        d = lambda a, b: r[a] - r[b]
        print(f"  stage {si:2d}: [{d(8,0):5d} {d(9,8):5d} {d(10,9):5d} {d(1,10):5d}] {d(2,1):5d} {d(3,2):5d} {d(4,3):5d} {d(5,4):5d} [{d(11,5):5d} {d(12,11):5d} {d(6,12):5d}] {d(7,6):5d}   total {d(7,0)}")
