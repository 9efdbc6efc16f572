"""Summarise an `ncu --set full` report (works without a GPU): per launch the duration, grid, DRAM bytes, pipe utilisation
and top stall reasons.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    def g(k, scale=None):
        v = num(d.get(k, ""))
        if v is None: return None
        unit = u.get(k, "")
        if unit in ("Kbyte", "KB"): v *= 1e3
        if unit in ("Mbyte", "MB"): v *= 1e6
        if unit in ("Gbyte", "GB"): v *= 1e9
        if unit == "ms": v *= 1e3
        if unit == "ns": v /= 1e3
        return v
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): num(d[k])
              for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and num(d[k])}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
    res.append({
        "kernel": d.get("Kernel Name", "")[:70], "grid": d.get("Grid Size"), "block": d.get("Block Size"),
        "duration_us": g("gpu__time_duration.sum"),
        "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
        "registers": g("launch__registers_per_thread"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_pct": g("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active") or g("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active") or g("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_realtime_pct": g("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        "lts_throughput_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "top_stalls": top,
    })
for r in res:
    print(json.dumps(r))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
