"""Split a kernel's SASS (from an .ncu-rep with source counters) into regions delimited by
BAR.SYNC instructions and print, per region: warp instructions executed, samples and the stall
reasons.  Development aid:  ncu_regions.py report.ncu-rep [min_exe]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ins.append(r)
seen, uniq = set(), []
for r in ins:
    if r[0] not in seen:
        seen.add(r[0]); uniq.append(r)
iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
iT = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
regions = []
cur = {"start": 0, "n": 0, "exe": 0, "smp": 0, "st": collections.Counter(), "first": None, "ops": collections.Counter(), "texe": 0}
for k, r in enumerate(uniq):
    e = int(r[iE] or 0); s = int(r[iS] or 0)
    cur["n"] += 1; cur["exe"] += e; cur["smp"] += s
    if iT is not None: cur["texe"] += int(r[iT] or 0)
    op = r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]
    cur["ops"][op.split(".")[0]] += e
    for i in stall:
        try: cur["st"][hdr[i][6:]] += int(r[i])
        except ValueError: pass
    if "BAR.SYNC" in r[1] or k == len(uniq) - 1:
        cur["end"] = k
        regions.append(cur)
        cur = {"start": k + 1, "n": 0, "exe": 0, "smp": 0, "st": collections.Counter(), "ops": collections.Counter(), "texe": 0}
tot_e = sum(x["exe"] for x in regions); tot_s = sum(x["smp"] for x in regions)
print(f"total warp-instr {tot_e}  samples {tot_s}")
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
for x in regions:
    if x["exe"] < mn and x["smp"] < 30: continue
    print(f"--- SASS [{x['start']}..{x['end']}] ({x['n']} instr): exe {x['exe']} ({100*x['exe']/tot_e:.1f}%)  samples {x['smp']} ({100*x['smp']/max(tot_s,1):.1f}%)"
          + (f"  active lanes/instr {x['texe']/max(x['exe'],1):.1f}" if x['texe'] else ""))
    print("    stalls: " + ", ".join(f"{k}={v}" for k, v in x["st"].most_common(9)))
    print("    ops:    " + ", ".join(f"{k}={v}" for k, v in x["ops"].most_common(14)))
