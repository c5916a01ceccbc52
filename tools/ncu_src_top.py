"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (compile with -lineinfo).
Usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_src_top.py src.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = r; body = rows[i + 1:]; break
if hdr is None:
    print("no source table; columns seen:", rows[0][:10]); sys.exit(1)
ci = {c: k for k, c in enumerate(hdr)}
samp = next(c for c in hdr if c.startswith("# Samples") or c == "Sampling Data (All)" or "Samples" in c)
top = []
for r in body:
    if len(r) != len(hdr): continue
    try: v = float(r[ci[samp]].replace(",", "") or 0)
    except ValueError: continue
    top.append((v, r[ci.get("#", 0)] if "#" in ci else "", r[ci["Source"]].strip()[:150]))
tot = sum(v for v, _, _ in top) or 1
print("column:", samp, "total samples", tot)
for v, ln, src in sorted(top, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  {ln:>5}  {src}")
