"""GPU: cProfile of the host side of 300 training steps (where does the enqueue time go?)."""
import cProfile, pstats, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import Config, PAMNet, ops
from pamnet_b200.data import synthetic_qm9_batch
torch.manual_seed(0)
model = PAMNet(Config("QM9", 128, 6, 5.0, 5.0)).cuda()
b = synthetic_qm9_batch(32, 0).to("cuda")
def step():
    model.zero_grad()
    out = model(b)
    loss = ops.l1_loss(out, b.y)
    loss.backward()
    model.prefetch(b, wait_current=False)
for _ in range(10): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300): step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("tottime")
st.print_stats(22)
