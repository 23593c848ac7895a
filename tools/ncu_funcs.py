"""Per-function / per-line instruction shares of one kernel from an ncu report (needs -lineinfo and --import-source on).
    python tools/ncu_funcs.py gpurun_out/prof.ncu-rep [units] [file:first-last]
units: divide the instruction counts by this number (e.g. warps x trades) to print instructions per unit."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
detail = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
d = dict(zip(hdr, rows[2]))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
          "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"):
    print("%-60s %s" % (k, d.get(k)))
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(d[k]) > 0.05:
        print("%-60s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), d[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
tot = collections.Counter(); samp = collections.Counter(); text = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]; continue
    if len(r) < 10 or r[0] in ("Line No", ""):
        continue
    try:
        ln = int(r[0]); ie = int(r[7]); sm = int(r[6])
    except ValueError:
        continue
    tot[(cur, ln)] += ie; samp[(cur, ln)] += sm; text[(cur, ln)] = r[1]
T = sum(tot.values()) or 1
S = sum(samp.values()) or 1
print("instructions %d (%.0f per unit), samples %d" % (T, T / units, S))
for f in sorted(set(k[0] for k in tot)):
    v = sum(x for (ff, l), x in tot.items() if ff == f)
    s = sum(x for (ff, l), x in samp.items() if ff == f)
    print("%-70s %5.1f%% inst %5.1f%% samples %8.0f/unit" % (os.path.basename(f), 100 * v / T, 100 * s / S, v / units))
    if not os.path.exists(f) or v < 0.02 * T:
        continue
    lines = open(f).read().split("\n")
    starts = [(i + 1, l) for i, l in enumerate(lines) if re.match(r"^(LN_HD|static|template|__global__|__device__|inline)", l)]
    starts.append((len(lines) + 1, "end"))
    for (a, name), (b, _) in zip(starts, starts[1:]):
        v = sum(x for (ff, l), x in tot.items() if ff == f and a <= l < b)
        s = sum(x for (ff, l), x in samp.items() if ff == f and a <= l < b)
        if v > 0.003 * T:
            print("    %-66s %5.1f%% inst %5.1f%% samples %8.0f/unit" % (name.strip()[:66], 100 * v / T, 100 * s / S, v / units))
if detail:
    fn, rng = detail.split(":")
    a, b = (int(x) for x in rng.split("-"))
    for k in sorted(tot):
        if os.path.basename(k[0]) == fn and a <= k[1] <= b and tot[k] > 0.001 * T:
            print("%5d %8.1f/unit %5.1f%% samp  %s" % (k[1], tot[k] / units, 100 * samp[k] / S, text[k][:110]))
