"""Warp-stall hot spots of one `ncu --set full --import-source on` capture: the SASS-level source page
(`ncu -i X.ncu-rep --page source --csv > src.csv`) aggregated by stall reason and the instructions holding the most samples.
    python scripts/ncu_hotspots.py src.csv > profiles/r02_rollout_tc2_source_hotspots.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
print("kernel:", rows[0][1])
head = rows[1]
data = [r for r in rows[2:] if len(r) == len(head)]
ci = {n: i for i, n in enumerate(head)}
stalls = [n for n in head if n.startswith("stall_") and "Not Issued" not in n]
num = lambda r, k: int(r[ci[k]] or 0)
tot = sum(num(r, "# Samples") for r in data)
print("warp-stall samples: %d over %d SASS instructions" % (tot, len(data)))
agg = sorted(((sum(num(r, s) for r in data), s) for s in stalls), reverse=True)
print("by reason:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * n / tot) for n, s in agg if n * 200 > tot))
print("top instructions (share of samples, SASS, dominant reasons):")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:30]:
    top = sorted(((num(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print("  %5.1f%%  %-72s %s" % (100.0 * num(r, "# Samples") / tot, r[ci["Source"]].strip()[:72],
                                   " ".join("%s=%d" % (s, n) for n, s in top if n)))
