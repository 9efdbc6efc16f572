"""GPU: warm-cache timing of single GEMM launches through the C ABI (CUDA events).  python tools/gemm_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import _lib

lib = _lib.load()
torch.manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
shapes = [  # mode, M, N, K, ksplit
    (0, 11346, 128, 128, 1), (0, 68076, 128, 128, 1), (0, 11346, 128, 16, 1), (0, 9992, 128, 88, 1),
    (1, 11346, 128, 128, 1), (1, 11346, 128, 768, 1), (1, 68076, 128, 128, 1),
    (2, 128, 128, 620, 5), (2, 128, 128, 11346, 89), (2, 128, 128, 68076, 532), (2, 128, 128, 128, 1),
]
for mode, M, N, K, ks in shapes:
    if mode == 0:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    elif mode == 1:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    else:
        a, b = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
    c = torch.zeros(M, N, device="cuda")
    def run():
        _lib.check(lib.pamnet_gemm(mode, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], c.data_ptr(), N, M, N, K,
                                   ks, None, st), "gemm")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"mode {mode} M {M} N {N} K {K} ksplit {ks}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.2f} TFLOP/s (fp32-equivalent)", flush=True)
