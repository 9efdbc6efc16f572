#!/bin/bash
# A/B of the build-time experiment PAMNET_FAST_SILU (ex2.approx + rcp.approx in SiLU, csrc/common.cuh) on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/fast_silu_ab.sh'
# Rebuilds the in-tree library with the flag (nvcc is on the box), runs the parity suite and the bench, then restores the
# default build.  Keep the flag only if gpurun_out/fast_silu_parity.log is green (the ladder of tests/helpers.py).
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/fast_silu_bench_default.json 2> gpurun_out/fast_silu.err
PAMNET_FAST_SILU=1 python physics-aware-multiplex-gnn_b200/build.py --force > gpurun_out/fast_silu_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/fast_silu_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/fast_silu_parity.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/fast_silu_bench_fast.json 2>> gpurun_out/fast_silu.err
python physics-aware-multiplex-gnn_b200/build.py --force >> gpurun_out/fast_silu_build.log 2>&1
tail -3 gpurun_out/fast_silu_parity.log
python - <<'PY'
import json
for f in ("default", "fast"):
    try:
        d = json.loads(open(f"gpurun_out/fast_silu_bench_{f}.json").read().strip().splitlines()[-1])
        k = d.get("kernels", {})
        print(f, "ms/step %.3f" % d["ms_per_step"], {n: round(v["ms_per_step"], 3) for n, v in k.items() if "msg" in n or n == "node_chain"})
    except Exception as exc:
        print(f, "unreadable:", exc)
PY
