import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import ops
torch.manual_seed(0)
m = n = 128
for mode in (1, 2):
    for k, lo, hi in [(16, 0, 8), (16, 8, 16), (16, 0, 16), (32, 0, 32)]:
        a = torch.zeros(m, k); b = torch.zeros(n, k)
        a[:, lo:hi] = 1.0
        b[:, lo:hi] = (torch.arange(n).float()[:, None] + 1)
        if mode == 1:
            A, B = a, b.T.contiguous()
        else:
            A, B = a.T.contiguous(), b.T.contiguous()
        ref = a @ b.T
        out = ops.gemm(mode, A.cuda(), B.cuda(), m, n, k).cpu()
        print("mode", mode, "K", k, "nonzero k in", (lo, hi), "out[0,:3]", out[0, :3].tolist(), "ref", ref[0, :3].tolist(), "nnz", int((out != 0).sum()), "out[1,:3]", out[1, :3].tolist(), "absmax", out.abs().max().item())
