"""Per-CUDA-source-line instruction / stall-sample table from an ncu report (needs -lineinfo, --import-source on):
    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep 'regex:spatialKernel' [top] [function-name substring]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sub = sys.argv[4] if len(sys.argv) > 4 else ""
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
agg = []
cur = "?"
keep = True
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        keep = sub in r[1]
        continue
    if r and r[0] == "Line No":
        hdr = {}
        for i, n in enumerate(r):
            hdr.setdefault(n, i)
        continue
    if hdr is None or not keep or len(r) < 10 or r[2] != "-":
        continue
    try:
        agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0),
                    int(r[hdr["Thread Instructions Executed"]] or 0)))
    except ValueError:
        pass
ti = sum(a[4] for a in agg) or 1
ts = sum(a[3] for a in agg) or 1
print("warp instr %d  samples %d  lanes/inst %.1f" % (ti, ts, sum(a[5] for a in agg) / ti))
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print("%-18s %4d inst %5.1f%% samp %5.1f%% lanes %4.1f | %s" % (a[0][:18], a[1], 100 * a[4] / ti, 100 * a[3] / ts, a[5] / max(a[4], 1), a[2]))
