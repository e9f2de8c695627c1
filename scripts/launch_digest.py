"""Digest of an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the number
of launches, total and mean duration and the share of the captured GPU time.
    python scripts/launch_digest.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4]
    short = name.split("<")[0].split("(")[0].replace("void ", "").strip()
    if "k_fixpoint_spec" in name:
        short = name.split("(")[0].replace("void ", "")
    agg[short][0] += 1
    agg[short][1] += float(r[-1].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us of GPU time (ncu time base, cold caches, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * v[1] / tot:6.2f} %  {v[0]:4d} launches  total {v[1]:10.1f} us  mean {v[1] / v[0]:9.1f} us  {k[:110]}")
