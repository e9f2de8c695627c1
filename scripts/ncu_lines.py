"""Aggregate an .ncu-rep's SASS samples / executed instructions per CUDA source line (development aid)."""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0, 0])
src = {}
hdr = None; fname = None
for r in rows:
    if r and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        k = (fname, int(r[0]))
        src[k] = r[1]
        try:
            agg[k][0] += int(r[iS] or 0); agg[k][1] += int(r[iE] or 0)
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()); tot_e = sum(v[1] for v in agg.values())
print("total samples", tot_s, "warp-instr", tot_e)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = (lambda kv: -kv[1][1]) if len(sys.argv) > 3 and sys.argv[3] == "exe" else (lambda kv: -kv[1][0])
for k, v in sorted(agg.items(), key=key)[:top]:
    print(f"{v[0]:6d} smp {100*v[0]/max(tot_s,1):5.1f}%  {v[1]:9d} exe {100*v[1]/max(tot_e,1):5.1f}%  {k[0]}:{k[1]:4d}  {src[k][:90]}")
