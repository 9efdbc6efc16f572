"""Per-source-line warp-stall samples of one kernel launch in an `ncu --set full --import-source on` report.

    python tools/ncu_source_hotspots.py gpurun_out/r01aq_full.ncu-rep [launch-skip] [top-n]

Reads the report through `ncu -i ... --page source --csv --print-source sass,cuda` (works without a GPU) and prints the
source lines ordered by their share of the samples -- the table committed as profiles/r01_source_hotspots.txt."""
import collections
import csv
import io
import subprocess
import sys


def main(rep, skip=0, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = next((r[1] for r in rows if r and r[0] in ("Function Name", "Kernel Name")), "?")
    hdr = next(r for r in rows if r and r[0] == "Line No")
    samp = hdr.index("# Samples")
    agg, text = collections.Counter(), {}
    for r in rows:
        if r and r[0].isdigit() and len(r) > samp and r[samp].isdigit():
            agg[int(r[0])] += int(r[samp])
            text[int(r[0])] = r[1].strip()
    tot = sum(agg.values()) or 1
    print(f"{name}  (launch {skip} of {rep}; {tot} samples)")
    for line, n in agg.most_common(top):
        print(f"{n:6d} {100.0 * n / tot:5.1f} %  line {line:4d}  {text[line][:110]}")


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], int(a[1]) if len(a) > 1 else 0, int(a[2]) if len(a) > 2 else 25)
