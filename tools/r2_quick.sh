#!/bin/bash
# quick parity + bench: usage bash tools/r2_quick.sh <tag> [pytest-args]
set -u
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu ${2:-} > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
PAMNET_STREAMS=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_1s.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batch-size 256 > gpurun_out/${tag}_bench256.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d.get("kernels", {})
        print(f, "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
        print("    ", {n: round(v["ms_per_step"], 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"])})
    except Exception as exc:
        print(f, "unreadable:", exc)
PY
tail -5 gpurun_out/${tag}_bench.err
