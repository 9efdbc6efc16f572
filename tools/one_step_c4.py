"""GPU: N training steps of BASELINE configs[3] (the 8-graph RNA fixture batch) -- workload for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet
from pamnet_b200.data import Batch
gold = torch.load(os.path.join(ROOT, "tests", "golden", "rna_c4.pt"), map_location="cpu", weights_only=False)
model = PAMNet(Config(**gold["config"])); model.load_state_dict(gold["state_dict"]); model = model.cuda()
b = Batch(x=gold["x"], batch=torch.repeat_interleave(torch.arange(8), torch.tensor(gold["sizes"])), y=gold["y"]).to("cuda")
for _ in range(int(os.environ.get("STEPS", "3"))):
    model.zero_grad()
    out = model(b)
    (out - b.y).abs().mean().backward()
    torch.cuda.synchronize()
print("done")
