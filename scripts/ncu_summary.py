"""Summarise an ncu --set full report into a small table for profiles/ (run on the CPU box):
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_summary.md
Also writes profiles/traffic.json (dram bytes per launch per kernel, read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
     ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
     ("smsp__inst_executed.sum", "warp_inst"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM%"),
     ("launch__registers_per_thread", "regs"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
     ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar")]


def scale(v, u):
    v = float(v.replace(",", ""))
    f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u)
    return v * f if f else v


agg = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void abx::<unnamed>::", "").replace("abx::<unnamed>::", "").strip()
    a = agg.setdefault(name, {"n": 0})
    a["n"] += 1
    for m, short in M:
        if m in col and r[col[m]] not in ("", "n/a"):
            try:
                a[short] = a.get(short, 0.0) + scale(r[col[m]], units[col[m]])
            except ValueError:
                pass
lines = ["# ncu --set full --clock-control none summary of `%s` (per-launch averages)" % os.path.basename(rep), "",
         "| kernel | launches | time ms | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 hit % | L1 hit % | occ % | lanes/inst | "
         "warp inst M | SM % | regs | stall long_sb | short_sb | barrier | membar |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
tj = os.path.join(os.path.dirname(out), "traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}  # merged: a partial profile only updates its kernels
seen = set()
for name, a in sorted(agg.items(), key=lambda kv: -kv[1].get("time", 0)):
    n = a["n"]
    g = lambda k: a.get(k, 0.0) / n
    t = g("time")
    bw = (g("dram_rd") + g("dram_wr")) / t / 1e9 if t else 0
    # keys: full name, and the bare kernel name for its longest-running instantiation (bench.py looks that up)
    clean = name.replace("void ", "").replace("unnamed>::", "").strip()
    traffic[clean] = g("dram_rd") + g("dram_wr")
    if clean.split("<")[0] not in seen:
        seen.add(clean.split("<")[0])
        traffic[clean.split("<")[0]] = g("dram_rd") + g("dram_wr")
    lines.append("| %s | %d | %.4f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.2f | %.2f | %.2f | %.2f |"
                 % (name, n, t * 1e3, g("dram_rd") / 1e6, g("dram_wr") / 1e6, bw, g("L2hit%"), g("L1hit%"), g("occ%"),
                    g("lanes/inst"), g("warp_inst") / 1e6, g("SM%"), g("regs"), g("st_long_sb"), g("st_short_sb"),
                    g("st_barrier"), g("st_membar")))
open(out, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(tj, "w"), indent=1)
print("\n".join(lines))
