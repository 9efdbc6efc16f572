"""GPU: N plain training steps (forward + L1 + backward) at the headline config -- the workload for ncu launch lists.
    BS=32 STEPS=3 python tools/one_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import synthetic_qm9_batch
torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(int(os.environ.get("BS", "32")), 0).to("cuda")
for _ in range(int(os.environ.get("STEPS", "3"))):
    for p in model.parameters():
        p.grad = None
    out = model(b)
    (out - b.y).abs().mean().backward()
    torch.cuda.synchronize()
print("done")
