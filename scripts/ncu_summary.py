"""Summarise an .ncu-rep: headline metrics per kernel + stall reasons + hottest SASS lines."""
import collections, csv, subprocess, sys, io

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('===', r[hdr.index('Kernel Name')][:70])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:62s} {r[i]:>16s} {units[i]}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "sass"]))))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ins.append(r)
# de-duplicate (ncu prints each line once per view)
seen, uniq = set(), []
for r in ins:
    k = (r[0], r[1])
    if k not in seen:
        seen.add(k); uniq.append(r)
iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in uniq:
    for i in stall:
        try: tot[hdr[i]] += int(r[i])
        except ValueError: pass
print("SASS instructions:", len(uniq), " samples:", sum(int(r[iS] or 0) for r in uniq),
      " warp-instr executed:", sum(int(r[iE] or 0) for r in uniq))
print("stalls:", ", ".join(f"{k[6:]}={v}" for k, v in tot.most_common(10)))
for r in sorted(uniq, key=lambda r: -int(r[iS] or 0))[:top]:
    print(f"  {r[iS]:>7s} smp {r[iE]:>9s} exe  {r[0][-5:]}  {r[1][:100]}")
