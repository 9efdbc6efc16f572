"""GPU: clock64 timeline of the node chain kernel (CTA 0) for the LAST chain launch of a training step.
    PAMNET_TC_TRACE=1 python physics-aware-multiplex-gnn_b200/build.py --force && python tools/chain_trace.py
Tensor-core interpreter (default): 8 stamps per stage -- start | fields + operand prefetch + prologue | weight wait |
barrier | multiply | epilogue math | stores + slot writes | trailing barrier."""
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
b = synthetic_qm9_batch(int(os.environ.get("BS", "32")), 0).to("cuda")
for which in ("forward", "backward"):
    for _ in range(3):
        for p in model.parameters():
            p.grad = None
        out = model(b)
        if which == "backward":
            (out - b.y).abs().mean().backward()
        torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 208)()
    _lib.check(lib.pamnet_debug_chain_trace(buf, 208), "trace")
    t = list(buf)
    print(f"last chain launch of {which}: per stage: pre wait bar mma epi store endbar | total")
    t0 = None
    for si in range(26):
        r = t[si * 8:(si + 1) * 8]
        if r[0] == 0 or r[7] == 0 or r[7] < r[0]:
            continue
        t0 = t0 or r[0]
        if r[1] > r[0] and r[6] > r[0]:
            d = lambda a, b: r[a] - r[b]
            print(f"  stage {si:2d} gemm: {d(1,0):5d} {d(2,1):5d} {d(3,2):5d} {d(4,3):5d} {d(5,4):5d} {d(6,5):5d} {d(7,6):5d} | {d(7,0):6d}   (t={r[0]-t0})")
        else:
            print(f"  stage {si:2d} other: {r[7]-r[0]:6d}   (t={r[0]-t0})")
