#!/bin/bash
# Round-2 A/B helper: parity suite then bench variants.  usage: bash tools/r2_ab.sh <tag> [extra env assignments for variant B]
set -u
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_$rep.json 2> gpurun_out/${tag}_bench.err
  PAMNET_CHAIN=ffma timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_ffma_$rep.json 2>> gpurun_out/${tag}_bench.err
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batch-size 256 > gpurun_out/${tag}_bench256.json 2>> gpurun_out/${tag}_bench.err
PAMNET_CHAIN=ffma timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batch-size 256 > gpurun_out/${tag}_bench256_ffma.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d.get("kernels", {})
        print(f, "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
              {n: round(v["ms_per_step"], 3) for n, v in k.items() if n in ("node_chain", "gemm_f32")})
    except Exception as exc:
        print(f, "unreadable:", exc)
PY
tail -5 gpurun_out/${tag}_bench.err
