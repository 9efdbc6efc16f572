#!/bin/bash
# serialized launch lists (both chain interpreters) + one full capture of the tensor-core chain kernel
set -u
tag=${1:-r2b}
mkdir -p gpurun_out
for mode in mma ffma; do
  PAMNET_CHAIN=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$mode.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/${tag}_ncu_$mode.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_mma -s 60 -c 3 -o gpurun_out/${tag}_chain_full \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/${tag}_ncu_full.log 2>&1
python - <<PY
import csv, collections
for mode in ("mma", "ffma"):
    rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches_%s.csv" % mode)) if len(r) > 5]
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except Exception: continue
        if d.get("Metric Unit") == "ns": v /= 1e3
        elif d.get("Metric Unit") == "ms": v *= 1e3
        name = d["Kernel Name"].split("(")[0][:60]
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(mode, "total us", round(tot))
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
        print("   %-60s %4d  %9.1f us  avg %7.1f" % (n, v[0], v[1], v[1] / v[0]))
PY
