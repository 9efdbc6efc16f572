"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST step's kernels by name (+ grid).
    python tools/launch_list.py gpurun_out/x.csv [n_steps]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr = None
recs = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except Exception: continue
    if d.get("Metric Unit") == "ns": v /= 1e3
    elif d.get("Metric Unit") == "ms": v *= 1e3
    recs.append((d["Kernel Name"].split("(")[0][-58:], d.get("Grid Size", ""), v))
pam = [r for r in recs if "pamnet" in r[0]]
marks = [i for i, r in enumerate(pam) if "loss_kernel" in r[0]]
if len(marks) >= 3 and len(sys.argv) <= 2:
    # one full cycle between two loss launches (backward of step i, plan + forward of step i + 1); the middle of the run
    a, b = marks[len(marks) // 2], marks[len(marks) // 2 + 1]
    last = pam[a:b]
    n = b - a
else:
    n = len(pam) // steps
    last = pam[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, grid, v in last:
    agg[(name, grid if "gemm" in name else "")][0] += 1
    agg[(name, grid if "gemm" in name else "")][1] += v
tot = sum(v[1] for v in agg.values())
print("one step: %d launches, %.0f us of kernel time" % (n, tot))
for (name, grid), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("  %-58s %-18s %3d  %8.1f us  avg %6.1f" % (name, grid, v[0], v[1], v[1] / v[0]))
