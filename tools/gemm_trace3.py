import ctypes, os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for mode, M, N, K, ks in [(0, 93540, 128, 128, 1), (1, 93540, 128, 128, 1), (0, 11346, 128, 128, 1)]:
    if mode == 0: a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    elif mode == 1: a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    c = torch.zeros(M, N, device="cuda")
    for _ in range(3):
        _lib.check(lib.pamnet_gemm(mode, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], c.data_ptr(), N, M, N, K, ks, None, st), "gemm")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(lib.pamnet_gemm(mode, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], c.data_ptr(), N, M, N, K, ks, None, st), "gemm")
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"mode {mode} M {M}: {us:.1f} us/launch  -> {(M*K*4 + M*N*4)/us/1e3:.0f} GB/s  {2*M*N*K/us/1e6:.1f} fp32-TFLOP/s")
    buf = (ctypes.c_longlong * 96)()
    _lib.check(lib.pamnet_debug_tc_trace(buf, 96), "trace")
    t = list(buf); t0 = t[0]
    print(f"  setup {t[1]-t0}  pdl {t[2]-t0}")
    for c_ in range(12):
        r = [t[8 + 4 * c_ + i] - t0 for i in range(4)]
        if r[0] < 0 or r[0] > 10**7: break
        print(f"  chunk {c_:2d}: tma issued {r[0]:6d}  data seen {r[1]:6d}  converted {r[2]:6d}  mma issued {r[3]:6d}")
    for i in range(4):
        r = [t[64 + 4 * i + j] - t0 for j in range(3)]
        if r[0] < 0 or r[0] > 10**7: break
        print(f"  item {i}: acc full {r[0]:6d}  tmem drained {r[1]:6d}  stores done {r[2]:6d}")
