"""Build libpamnet_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpamnet_sm100.so")
SOURCES = ["abi.cu", "comm.cu", "graph.cu", "graph_grid.cu", "front_mol.cu", "collate.cu", "basis.cu", "gemm.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_small.cu", "chain.cu", "chain_mma.cu", "message.cu", "readout.cu", "model.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
if os.environ.get("PAMNET_SILU_MODE"):           # 0 exact / 1 fast reciprocal (default) / 2 + ex2.approx (csrc/common.cuh)
    NVCC_FLAGS.append("-DPAMNET_SILU_MODE=" + str(int(os.environ["PAMNET_SILU_MODE"])))
if os.environ.get("PAMNET_TC_TRACE"):          # in-kernel clock64 timeline of the tensor-core GEMM (debug builds)
    NVCC_FLAGS.append("-DPAMNET_TC_TRACE")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pamnet_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or os.path.getmtime(o) < hdr_time:
            jobs.append([_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
