"""Top SASS instructions by stall samples from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python scripts/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
data = [r for r in rows[hdr_i + 1:end] if len(r) >= len(hdr)]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:N]
for i in sorted(order):
    r = data[i]
    st = {hdr[j][6:]: int(r[j]) for j in stall_cols if int(r[j] or 0)}
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}% x{r[ie]:>9s} {r[src].strip()[:70]:70s} {st}")
