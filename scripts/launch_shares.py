import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki][:60], [0, 0.0]); a[0] += 1; a[1] += float(r[vi]) / 1000
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:62s} {n:4d} {t:9.1f} {t / n:8.1f} {100 * t / tot:5.1f}%")
