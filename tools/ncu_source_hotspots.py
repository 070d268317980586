"""top source lines of an `ncu --page source --csv` export by warp-stall samples: usage ncu_source_hotspots.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr, body = r, rows[i + 1:]
        break
if hdr is None:
    print("no source table found; header candidates:", rows[:3])
    sys.exit(0)
si = hdr.index("Source")
samp = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or c == "Warp Stall Sampling (All Cycles)"]
inst = [i for i, c in enumerate(hdr) if c == "Instructions Executed"]
ci = samp[0] if samp else None
print("columns used:", hdr[ci] if ci is not None else None, "|", hdr[inst[0]] if inst else None)
lines = []
for r in body:
    if len(r) <= si or ci is None:
        continue
    try:
        v = float(r[ci].replace(",", "") or 0)
    except ValueError:
        continue
    ie = r[inst[0]] if inst else ""
    lines.append((v, ie, r[si].strip()[:150]))
total = sum(v for v, _, _ in lines) or 1.0
for v, ie, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * v / total:6.2f}%  samples={int(v):8d}  inst={ie:>12s}  {src}")
