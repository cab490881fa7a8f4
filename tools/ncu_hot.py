"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output (one kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {n:7d} {100.0 * n / tot:5.1f}%  exec={r[ix['Instructions Executed']]:>9}  {r[ix['Source']].strip()[:90]:90s} {st}")
