import csv, json, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
byt = json.load(open(sys.argv[2]))
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6547.8
best = collections.defaultdict(list)
for r in rows[1:]:
    best[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
print(f"standalone HBM-bound ops at 160,000 rays (BASELINE configs[1] sizes), ncu gpu__time_duration per launch (median of the flushed launches), peak {peak:.0f} GB/s")
for key, nbytes in byt.items():
    for kname, ts in best.items():
        if key in kname and "tile" not in key:
            ts = sorted(ts[-3:])
            t = ts[len(ts) // 2]
            print(f"  {kname:42s} {nbytes/1e6:8.1f} MB  {t/1e3:8.1f} us  {nbytes/t:8.1f} GB/s  {nbytes/t/peak:5.3f} of peak")
