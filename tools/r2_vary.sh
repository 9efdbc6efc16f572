#!/bin/bash
# fixed batch vs an epoch-like stream of distinct batches, configs c2 / c3 / c4
set -u
mkdir -p gpurun_out
tag=${1:-r2v}
run() { name=$1; shift; timeout 400 python bench.py --no-cpu-baseline "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run c2 --steps 100 --warmup 5
run c2_vary --steps 100 --warmup 5 --vary 16 --no-profile
run c3 --steps 30 --warmup 5 --config c3
run c3_vary --steps 30 --warmup 5 --config c3 --vary 8 --no-profile
run c4 --steps 50 --warmup 5 --config c4
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("  ", f.split("/")[-1], "ms/step %.3f" % d["ms_per_step"], d["ms_per_step_stats"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    except Exception as exc:
        print("  ", f, "unreadable:", exc)
PY
