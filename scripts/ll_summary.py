"""summary of an ncu launch list (scripts/launch_list.sh): per kernel / dynamic-smem class: count, mean, max, total duration"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]]})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
seq = []
for k, v in per.items():
    nm = v["name"].split("(")[0].split("<")[0]
    seq.append((nm, int(v.get("launch__shared_mem_per_block_dynamic", 0)), v.get("gpu__time_duration.sum", 0.0) / 1e3, int(v.get("launch__grid_size", 0))))
agg = collections.OrderedDict()
for nm, sm, us, g in seq:
    a = agg.setdefault((nm, sm, g), [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
for (nm, sm, g), a in agg.items():
    print(f"{nm:28s} smem {sm:7d} grid {g:6d}  n {a[0]:5d}  mean {a[1] / a[0]:9.1f} us  max {a[2]:9.1f} us  total {a[1] / 1e3:8.2f} ms ({100 * a[1] / tot:4.1f} %)")
print("total %.2f ms over %d launches" % (tot / 1e3, len(seq)))
if len(sys.argv) > 2:
    for nm, sm, us, g in seq[:int(sys.argv[2])]: print(f"  {nm:24s} {sm:7d} {us:9.1f}")
