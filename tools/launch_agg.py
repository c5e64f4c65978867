"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u == "ns" else v if u in ("us", "usecond") else v * 1e3 if u == "ms" else v
    a = agg.setdefault(r[ki][:60], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"# total kernel time {tot/1e3:.3f} ms over {len(rows)-1} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} n={n:5d} total={t/1e3:9.3f} ms share={100*t/tot:6.2f}% avg={t/n:9.2f} us")
