"""Per-SASS-instruction execution counts from `ncu -i rep --page source --csv` output:
python tools/ncu_sass_hot.py src.csv [min_millions]   -> compact listing + total."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
isrc, iex, ithr, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("# Samples")
tot = 0
for n, r in enumerate(rows[2:]):
    if len(r) <= iex or not r[iex]:
        continue
    ex = float(r[iex]); tot += ex
    if ex / 1e6 >= thr:
        print(f"{n:5d} {ex/1e6:8.2f}M thr {r[ithr]:>5s} smp {r[ismp]:>6s}  {r[isrc][:90]}")
print("total", tot / 1e6, "M")
