import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    if r[iu] in ("ns", "nsecond"): v /= 1e3
    elif r[iu] in ("ms", "msecond"): v *= 1e3
    name = r[ik].split("(")[0][:70]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:9.3f} ms {v[0]:6d}  {v[1]/v[0]:8.1f} us  {k}")
print("total ms", tot / 1e3)
