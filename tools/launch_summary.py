"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one ED step, kernel by kernel."""
import csv, re, sys
path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(r['Kernel Name'], r.get('Grid Size'), float(r['Metric Value']) / 1e3) for r in csv.DictReader(lines)]
names = [re.sub(r'urnn::', '', re.sub(r'\(.*', '', n)) for n, _, _ in rows]
start = None
for i in range(len(rows) - 2):
    if 'head_kernel<3>' in names[i] or 'head_stage_kernel<3' in names[i]:
        start = i + 1
        break
end = start
while end < len(rows) and 'head_kernel<3>' not in names[end] and 'head_stage_kernel<3' not in names[end]:
    end += 1
tot = 0.0
for i in range(start, end + 1):
    print(f"{i - start:3d} {names[i][:70]:70s} {rows[i][1]:>14s} {rows[i][2]:8.1f} us")
    tot += rows[i][2]
print(f"one step: {end + 1 - start} launches, {tot:.1f} us (serialised, cold-cache ncu timings)")
