"""GPU: accuracy and timing of the GEMM paths.  PAMNET_GEMM=ffma|tc python tools/gemm_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import ops, _lib

lib = _lib.load()
g = torch.Generator().manual_seed(0)
print("backend", os.environ.get("PAMNET_GEMM", "tc(default)"))
shapes = [(0, 11346, 128, 128, 1), (0, 300, 128, 128, 1), (0, 1000, 256, 64, 1), (1, 11346, 128, 1536, 1), (1, 1780, 128, 512, 1),
          (2, 128, 128, 11346, 64), (2, 128, 128, 620, 5), (2, 128, 384, 2000, 8), (0, 130, 100, 16, 1)]
for mode, m, n, k, ks in shapes:
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g)
    if mode == 0:
        A, B, ref = a, b, a.double() @ b.double().T
    elif mode == 1:
        A, B, ref = a, b.T.contiguous(), a.double() @ b.double().T
    else:
        A, B, ref = a.T.contiguous(), b.T.contiguous(), a.double() @ b.double().T
    Ac, Bc = A.cuda(), B.cuda()
    out = ops.gemm(mode, Ac, Bc, m, n, k, ksplit=ks)
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item()
    # timing
    c = torch.zeros((m, n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(lib.pamnet_gemm(mode, Ac.data_ptr(), Ac.shape[1], Bc.data_ptr(), Bc.shape[1], c.data_ptr(), n, m, n, k, ks, None, st))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"mode {mode} M={m:6d} N={n:4d} K={k:6d} ks={ks:2d}  max|err|={err:.3e} (|ref|max {ref.abs().max():.1f})  {us:8.1f} us  {2*m*n*k/us/1e6:7.2f} TFLOP/s")
