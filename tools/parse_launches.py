"""ncu --csv launch list -> per-launch text + per-kernel aggregate."""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src, errors="ignore")))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
L = collections.OrderedDict()
for d in data:
    e = L.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]})
    e[d["Metric Name"]] = d["Metric Value"]
out, tot = [], 0.0
for i, e in L.items():
    t = float(e["gpu__time_duration.sum"].replace(",", "")) / 1e3
    tot += t
    out.append((int(i), e["name"].split("(")[0][:48], t, e["grid"], e["block"]))
agg = collections.defaultdict(lambda: [0, 0.0])
for o in out:
    agg[o[1]][0] += 1
    agg[o[1]][1] += o[2]
with open(dst, "w") as f:
    f.write(f"# {len(out)} launches, {tot:.1f} us total (ncu gpu__time_duration: cold-cache, serialised)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1]:9.1f} us  {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  {k}\n")
    f.write("\n# per launch\n")
    for o in out:
        f.write(f"{o[0]:4d} {o[2]:9.1f} us grid={o[3]:>14s} blk={o[4]:>12s} {o[1]}\n")
print(open(dst).read().split("# per launch")[0])
