"""GPU: in-kernel clock64 timeline of one tensor-core GEMM CTA.  Build with PAMNET_TC_TRACE=1 first:
    PAMNET_TC_TRACE=1 python physics-aware-multiplex-gnn_b200/build.py --force && python tools/gemm_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pamnet_b200
from pamnet_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for mode, M, N, K, ks in [(0, 11346, 128, 128, 1), (0, 68076, 128, 128, 1), (2, 128, 128, 11346, 89), (1, 11346, 128, 128, 1)]:
    if mode == 0:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    elif mode == 1:
        a, b = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
    else:
        a, b = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
    c = torch.zeros(M, N, device="cuda")
    for _ in range(3):
        _lib.check(lib.pamnet_gemm(mode, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1], c.data_ptr(), N, M, N, K, ks, None, st), "gemm")
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 256)()
    _lib.check(lib.pamnet_debug_tc_trace(buf, 256), "trace")
    t = list(buf)
    t0 = t[0]
    nch = min(16, (K // ks + 15) // 16)
    print(f"mode {mode} M {M} K {K} ksplit {ks}: setup {t[1]-t0}  epilogue start {t[2]-t0}  end {t[3]-t0}")
    for kc in range(nch):
        conv = [t[16 + 4 * kc + i] - t0 for i in range(4)]
        mma = [t[128 + 2 * kc + i] - t0 for i in range(2)]
        print(f"  chunk {kc}: landed {conv[0]} free {conv[1]} converted {conv[2]} issued {conv[3]} | mma ready {mma[0]} issued {mma[1]}")
