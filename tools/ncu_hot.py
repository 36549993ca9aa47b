"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1], "total samples", tot)
agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
print("stall mix:", ", ".join(f"{h[6:]}={v * 100 // max(tot, 1)}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    main = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {s * 100.0 / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} {main[0][1]}={main[0][0]} {main[1][1]}={main[1][0]}")
