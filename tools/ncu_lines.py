"""per-source-line shares of executed warp instructions and stall samples from `ncu --page source --csv --print-source cuda,sass`
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv; python tools/ncu_lines.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) / 100 if len(sys.argv) > 2 else 0.004
cur, hdr, agg = None, None, {}
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0] != "":
        try: ln = int(r[0])
        except ValueError: continue
        a = agg.setdefault((cur, ln), [0, 0, r[1][:100]])
        a[0] += num(r[7]); a[1] += num(r[6])
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions", tot, "samples", ts)
for (f, ln), (i, s, src) in sorted(agg.items()):
    if i / tot > thr or s / ts > thr:
        print("%-16s %4d %5.1f%% inst %5.1f%% smp  %s" % (f, ln, 100 * i / tot, 100 * s / ts, src))
