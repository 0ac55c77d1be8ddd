"""ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel launches / total ms / share (sampler excluded)."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iV:
        continue
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    v = float(r[iV].replace(",", ""))
    ms = v / 1e6 if r[iU].startswith("n") else (v / 1e3 if r[iU].startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(v[1] for k, v in agg.items() if "gibbs" not in k)
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = "    nan" if "gibbs" in k else f"{v[1] / tot:7.3f}"
    print(f"{k[:70]:70s} {v[0]:8d} {v[1]:10.3f} {share}")
