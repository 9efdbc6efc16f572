#!/bin/bash
# bench variants: usage bash tools/r2_bench3.sh <tag> "<ENV1>" "<ENV2>" ...   (each ENV string is a set of assignments)
set -u
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_v${i}.json 2> gpurun_out/${tag}_v${i}.err
  env $envs PAMNET_STREAMS=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_v${i}_1s.json 2>> gpurun_out/${tag}_v${i}.err
  env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --batch-size 256 > gpurun_out/${tag}_v${i}_256.json 2>> gpurun_out/${tag}_v${i}.err
  echo "== v$i: $envs"
  python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_v${i}*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d.get("kernels", {})
        print("  ", f.split("/")[-1], "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
              {n: round(v["ms_per_step"], 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"])[:3]})
    except Exception as exc:
        print("  ", f, "unreadable:", exc)
PY
done
