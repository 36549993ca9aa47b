"""DRAM traffic of the bench's roofline kernel (the decoder stage-1 cell step = sweep A + sweep B + blend) from an
`ncu --set full` capture made with tools/prof_cell.py:  ncu -i cell.ncu-rep --page raw --csv | python tools/ncu_traffic.py"""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3}.get(u, 1)
out = {"kernels": [], "dram_bytes_per_cell_step": 0.0, "us_per_cell_step_under_ncu": 0.0}
for r in data:
    rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
    out["kernels"].append({"name": r[ix["Kernel Name"]][:60], "dram_read": rd, "dram_write": wr, "us": t,
                           "dram_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                           "tensor_pct": float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                           "issue_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]])})
    out["dram_bytes_per_cell_step"] += rd + wr
    out["us_per_cell_step_under_ncu"] += t
print(json.dumps(out, indent=1))
