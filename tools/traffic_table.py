"""DRAM traffic of every launch of one f16x3 time step, from an ncu pass with three metrics over tools/step_table.py:

    URNN_T=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/r2_traffic_step.csv python tools/step_table.py gpurun_out/names.json
    python tools/traffic_table.py gpurun_out/r2_traffic_step.csv gpurun_out/names.json > profiles/r2_traffic_cells.json

Launch names come from urnn_ed_profile_dev (names.json): a step is stem1, the program's launches in order, then the four
head kernels.  The LAST complete step of the capture is used (caches warm, weight images built)."""
import csv, json, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
names = [n for n, _ in json.load(open(sys.argv[2]))["ops"]]            # stem1, ..., head
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
launch = {}
for r in rows[1:]:
    lid = int(r[ix["ID"]])
    d = launch.setdefault(lid, {"kernel": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * mult.get(r[ix["Metric Unit"]], 1.0)
seq = [launch[k] for k in sorted(launch)]
starts = [i for i, l in enumerate(seq) if "stem1_kernel" in l["kernel"]]
nbody = len(names) - 2
s0 = starts[-1]
body = seq[s0:s0 + 1 + nbody + 4]
assert len(body) == nbody + 5 and "head_kernel" in body[1 + nbody]["kernel"], "capture does not end with a complete step"
per = {}
for i, l in enumerate(body):
    nm = names[i] if i <= nbody else "head"
    e = per.setdefault(nm, {"dram_read": 0.0, "dram_write": 0.0, "us_under_ncu": 0.0})
    e["dram_read"] += l["dram__bytes_read.sum"]; e["dram_write"] += l["dram__bytes_write.sum"]; e["us_under_ncu"] += l["gpu__time_duration.sum"]
def cell(prefix):
    return sum(v["dram_read"] + v["dram_write"] for k, v in per.items() if k.startswith(prefix + "."))
out = {"dec1": cell("dec1"), "enc1": cell("enc1"), "whole_step": sum(v["dram_read"] + v["dram_write"] for v in per.values()),
       "per_launch": {k: {a: round(b, 1) for a, b in v.items()} for k, v in per.items()},
       "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one warm step at 500x500 (tools/traffic_table.py)"}
print(json.dumps(out, indent=1))
