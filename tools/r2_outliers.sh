#!/bin/bash
# rare slow steps vs the clock sampler: none / in-process NVML / nvidia-smi child, 5 runs each
set -u
mkdir -p gpurun_out
for rep in 1 2 3 4 5; do
  for m in none nvml smi; do
    BENCH_SAMPLER=$m timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/r2o_${m}_$rep.json 2> gpurun_out/r2o_${m}_$rep.err
  done
done
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d["ms_per_step_stats"]
        print(f.split("/")[-1], "mean %.3f median %.3f max %.2f" % (d["ms_per_step"], st["median"], st["max"]), st["slow_steps"][:5], "e2e %.3f" % d["e2e"]["ms_per_step"], d["clocks"].get("samples"))
    except Exception as e:
        print(f, e)
PY
