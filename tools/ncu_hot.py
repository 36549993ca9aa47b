"""Top stalled SASS instructions per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
Usage: ncu_hot.py src.csv [top=30] [kernel-name-substring]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}; sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["body"].append(r)
for sec in sections:
    if want and want not in sec["name"]:
        continue
    hdr, body = sec["hdr"], sec["body"]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
    spin = sum(int(r[ix["Instructions Executed"]] or 0) for r in body if any(t in r[ix["Source"]] for t in ("TRYWAIT", "NANOSLEEP")))
    print("kernel:", sec["name"], "| total samples", tot)
    print(f"warp instructions executed {ex}, of which mbarrier polling (TRYWAIT/NANOSLEEP) {spin} = {spin * 100.0 / max(ex, 1):.1f}%")
    agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
    print("stall mix:", ", ".join(f"{h[6:]}={v * 100 // max(tot, 1)}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        s = int(r[ix["# Samples"]] or 0)
        main = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {s:6d} {s * 100.0 / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} {main[0][1]}={main[0][0]} {main[1][1]}={main[1][0]}")
    print()
