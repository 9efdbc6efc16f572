#!/bin/bash
# 3 vs 4 tensor-core products per k-step in the 3xTF32 GEMM: parity margin + bench (c2, c3)
set -u
mkdir -p gpurun_out
for p in 3 4; do
  export PAMNET_TC2_PROD=$p
  python tools/parity_margin.py 2>&1 | head -6
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_c2_$p.json 2> gpurun_out/r2p_$p.err
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --config c3 > gpurun_out/r2p_c3_$p.json 2>> gpurun_out/r2p_$p.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/r2p_c*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "ms/step %.3f" % d["ms_per_step"], d["ms_per_step_stats"], "gemm iso", round(d["roofline"]["class_ms_per_step_isolated"], 3))
PY
