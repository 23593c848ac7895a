"""Dynamic instructions per source function of one kernel, INCLUDING inline-PTX loops: joins the executed counts of the ncu
SASS page with the line table of the object file on the instruction offset.
    python tools/ncu_join.py gpurun_out/prof.ncu-rep superdendrix_b200/build/sdx_lanes.o '_Z13k_trade_lanesILb0ELi80E' superdendrix_b200/csrc/sdx_lane_trade.cuh [units]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, sym, header = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
hdr, rows = None, []
for r in csv.reader(io.StringIO(sass)):
    if r and r[0] == "Address":
        hdr = r
    elif len(r) > 10 and r[0].startswith("0x"):
        rows.append(r)
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[0][0], 16)
dyn = {int(r[0], 16) - base: (int(r[ie]), int(r[isamp])) for r in rows}
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + sym)][0]
cur, tot, samp = None, collections.Counter(), collections.Counter()
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and cur and int(m.group(1), 16) in dyn:
        tot[cur] += dyn[int(m.group(1), 16)][0]
        samp[cur] += dyn[int(m.group(1), 16)][1]
T, S = sum(tot.values()) or 1, sum(samp.values()) or 1
print("instructions %d (%.0f per unit)" % (T, T / units))
src = open(header).read().split("\n")
hb = os.path.basename(header)
starts = [(i + 1, l) for i, l in enumerate(src) if re.match(r"^(LN_HD|static|template|__device__|__global__)", l)] + [(len(src) + 1, "end")]
for (a, name), (b, _) in zip(starts, starts[1:]):
    v = sum(x for (f, l), x in tot.items() if f == hb and a <= l < b)
    s = sum(x for (f, l), x in samp.items() if f == hb and a <= l < b)
    if v > 0.003 * T:
        print("  %-66s %5.1f%% inst %5.1f%% samples %8.0f/unit" % (name.strip()[:66], 100 * v / T, 100 * s / S, v / units))
for f in sorted(set(k[0] for k in tot)):
    v = sum(x for (ff, l), x in tot.items() if ff == f)
    s = sum(x for (ff, l), x in samp.items() if ff == f)
    print("%-68s %5.1f%% inst %5.1f%% samples %8.0f/unit" % (f, 100 * v / T, 100 * s / S, v / units))
