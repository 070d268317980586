"""groups the SASS of an `ncu --page source --csv` export into regions of equal execution count: where the warp
instructions and the stall samples go.  usage: ncu_regions.py src.csv [listing.txt]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, body = rows[1], rows[2:]
si, ie, te, ss = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
seq = []
for r in body:
    if len(r) <= ie:
        continue
    try:
        seq.append((int(r[ie]), int(r[te]), int(r[ss]), r[si].strip()))
    except ValueError:
        continue
tot = sum(x[0] for x in seq)
tots = sum(x[2] for x in seq) or 1
print("total inst %.2fG static %d samples %d" % (tot / 1e9, len(seq), tots))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(f"{i + 1:5d} {n / 1e6:10.1f}M {t / max(n, 1):5.1f} {s:7d}  {src}"
                                           for i, (n, t, s, src) in enumerate(seq)))
groups, cur = [], [(0,) + seq[0]]
for i, r in enumerate(seq[1:], 1):
    if abs(r[0] - cur[-1][1]) <= 0.03 * max(cur[-1][1], 1):
        cur.append((i,) + r)
    else:
        groups.append(cur)
        cur = [(i,) + r]
groups.append(cur)
for g in groups:
    s = sum(r[1] for r in g)
    smp = sum(r[3] for r in g)
    if s / tot > 0.006 or smp / tots > 0.01:
        ops = {}
        for r in g:
            op = r[4].split()[0] if not r[4].startswith("@") else r[4].split()[1]
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
        print(f"lines {g[0][0] + 1:5d}-{g[-1][0] + 1:5d} n={len(g):4d} each={g[0][1] / 1e6:8.1f}M "
              f"thr={sum(r[2] / max(r[1], 1) for r in g) / len(g):5.1f} inst={100 * s / tot:5.1f}% "
              f"samples={100 * smp / tots:5.1f}% | {top}")
