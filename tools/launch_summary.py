import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr + 1:]
ki = cols.index('Kernel Name'); vi = cols.index('Metric Value'); ui = cols.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    if u == 'ns': v /= 1000
    elif u == 'ms': v *= 1000
    name = r[ki].split('(')[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f%% |" % (k[:60], n, t, t / n, 100 * t / tot))
