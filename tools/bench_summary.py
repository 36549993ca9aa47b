import json,sys
d=json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","dtype","clocks","n_gpus","scaling") if k in d})
if "e2e" in d: print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "dense", d["e2e_dense"]["value"], d["e2e_dense"]["ms_per_step"])
r=d.get("roofline")
if r: print("roofline dec1", r["frac"], r["ms_per_launch"], "enc1", r["worst_cell"]["frac"], "step", r["whole_step"]["frac"])
print(d.get("cpu_baseline")); print(d.get("sharded_parity"))
