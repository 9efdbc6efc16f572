"""GPU: timeline of one training step (per stream busy time, gaps, overlap).  python tools/timeline.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet, _lib
from pamnet_b200.data import synthetic_qm9_batch

torch.manual_seed(0)
if os.environ.get("CONFIG") == "c4":        # BASELINE configs[3]: the 8-graph RNA fixture batch
    from pamnet_b200.data import Batch
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "rna_c4.pt"), map_location="cpu", weights_only=False)
    model = PAMNet(Config(**gold["config"])); model.load_state_dict(gold["state_dict"]); model = model.cuda()
    b = Batch(x=gold["x"], batch=torch.repeat_interleave(torch.arange(8), torch.tensor(gold["sizes"])), y=gold["y"]).to("cuda")
else:
    model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
    b = synthetic_qm9_batch(int(os.environ.get("BS", "32")), 0).to("cuda")
params = list(model.parameters())

def step():
    for p in params:
        p.grad = None
    out = model(b)
    (out - b.y).abs().mean().backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
_lib.profile_begin()
step()
tl = _lib.profile_timeline()
print("launches", len(tl), "span %.3f ms" % (max(t[3] for t in tl) - min(t[2] for t in tl)))
for tag in sorted({t[1] for t in tl}):
    ev = sorted([t for t in tl if t[1] == tag], key=lambda t: t[2])
    busy = sum(t[3] - t[2] for t in ev)
    gaps = [(ev[i + 1][2] - ev[i][3], ev[i][0], ev[i + 1][0], ev[i][3]) for i in range(len(ev) - 1)]
    print(f"stream {tag}: {len(ev)} launches, busy {busy:.3f} ms, first start {ev[0][2]:.3f}, last end {ev[-1][3]:.3f}")
    big = sorted(gaps, reverse=True)[:8]
    print("   largest gaps (ms, after, before, at):", [(round(g[0], 3), g[1], g[2], round(g[3], 3)) for g in big])
    by = {}
    for t in ev:
        by.setdefault(t[0], [0, 0.0])
        by[t[0]][0] += 1; by[t[0]][1] += t[3] - t[2]
    print("   ", {k: (v[0], round(v[1], 3)) for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])})
# full trace, by start time
print("trace (stream:class @start ms +dur us):")
print("  ".join(f"{t[1]}:{t[0][:10]}@{t[2]:.3f}+{(t[3]-t[2])*1e3:.0f}" for t in sorted(tl, key=lambda t: t[2])))
