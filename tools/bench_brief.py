"""stdin: bench.py output -> one short summary (ms/step, stats, e2e, kernel classes)."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("ms/step %.3f" % d["ms_per_step"], d["ms_per_step_stats"], "e2e %.3f" % d["e2e"]["ms_per_step"], "launches", d.get("gpu_launches"))
print({k: round(v["ms_per_step"], 3) for k, v in d.get("kernels", {}).items()})
