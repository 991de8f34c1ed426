"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
src, title = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    name = r[ik].split("(")[0].replace("void cvo_b200::", "")
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
tot = sum(d[1] for d in agg.values())
print("#", title)
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
print(f"{'kernel':32s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:32]:32s} {d[0]:8d} {d[1]:12.1f} {d[1]/d[0]:10.1f} {100*d[1]/tot:6.1f}%")
print(f"{'total':32s} {sum(d[0] for d in agg.values()):8d} {tot:12.1f}")
