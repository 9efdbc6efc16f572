"""GPU: small diagnostic cases for the tensor-core GEMM paths (prints error and a corner of the output)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import ops
g = torch.Generator().manual_seed(0)
for mode, m, n, k, ks in [(1, 128, 128, 128, 1), (1, 256, 128, 64, 1), (2, 128, 128, 128, 1), (2, 128, 128, 32, 1), (2, 128, 128, 256, 2),
                          (1, 300, 128, 512, 1), (2, 256, 128, 1000, 4)]:
    a = torch.randn(m, k, generator=g); b = torch.randn(n, k, generator=g)
    if os.environ.get("ONES"):
        a = torch.ones(m, k); b = torch.arange(n).float()[:, None].expand(n, k).contiguous()
    ref = a.double() @ b.double().T
    if mode == 1: A, B = a, b.T.contiguous()
    else: A, B = a.T.contiguous(), b.T.contiguous()
    out = ops.gemm(mode, A.cuda(), B.cuda(), m, n, k, ksplit=ks)
    torch.cuda.synchronize()
    o = out.double().cpu()
    print(f"mode {mode} M={m} N={n} K={k} ks={ks}: max|err| {float((o-ref).abs().max()):.3e}  |ref|max {float(ref.abs().max()):.1f}  nonzero {int((o!=0).sum())}/{o.numel()}")
    print("   out[0,:6]", [round(float(x),3) for x in o[0,:6]], " ref[0,:6]", [round(float(x),3) for x in ref[0,:6]])
    print("   out[5,32:36]", [round(float(x),3) for x in o[5,32:36]], " ref", [round(float(x),3) for x in ref[5,32:36]])
