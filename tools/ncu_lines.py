"""Per-source-line instruction / stall-sample shares from an ncu report captured with --import-source on:
    ncu -i REPORT.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [N]"""
import csv, sys

def num(x):
    try:
        return int(x)
    except ValueError:
        return 0

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0] not in ("", "Function Name") and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        ie = num(d.get("Instructions Executed", "0"))
        if ie:
            data.append((cur, int(r[0]), r[1], ie, num(d["# Samples"]), num(d["Thread Instructions Executed"]), num(d.get("stall_barrier", "0"))))
tot = sum(x[3] for x in data)
ts = sum(x[4] for x in data)
print("total warp instructions", tot, "samples", ts)
byfile = {}
for f, l, s, ie, sm, ti, sb in data:
    a = byfile.setdefault(f, [0, 0])
    a[0] += ie
    a[1] += sm
for f, (a, b) in byfile.items():
    print("%-20s %5.1f%% instr %5.1f%% samples" % (f, 100 * a / tot, 100 * b / ts))
data.sort(key=lambda x: -x[3])
for f, l, s, ie, sm, ti, sb in data[:top]:
    print("%5.2f%% s%5.2f%% thr%5.1f %s:%d | %s" % (100 * ie / tot, 100 * sm / ts, ti / ie, f[:14], l, s.strip()[:100]))
