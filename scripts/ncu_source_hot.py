#!/usr/bin/env python
"""Per-instruction view of an .ncu-rep (source page, SASS): executed-instruction counts and stall samples,
aggregated over address ranges or printed for the hottest lines.
usage: ncu_source_hot.py rep [--ranges a-b,c-d (hex)] [--top N] [--dump a-b]"""
import csv, subprocess, sys, argparse, io
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--ranges", default=""); ap.add_argument("--top", type=int, default=0); ap.add_argument("--dump", default="")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
recs = []
for r in rows[2:]:
    if len(r) <= iexec: continue
    addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = addr
    st = {hdr[i]: int(r[i] or 0) for i in stall_cols}
    recs.append((addr - base, r[isrc], int(r[isamp] or 0), int(r[iexec] or 0), st))
tot_exec = sum(r[3] for r in recs); tot_samp = sum(r[2] for r in recs)
print(f"total executed {tot_exec/1e6:.1f} M warp-instr, samples {tot_samp}")
def is_fp64(s): return s.split()[0].replace("@","").startswith(("DADD","DMUL","DFMA")) or any(t in s for t in (" DADD "," DMUL "," DFMA "))
print(f"FP64 executed {sum(r[3] for r in recs if is_fp64(r[1]))/1e6:.1f} M")
for rg in [x for x in a.ranges.split(",") if x]:
    lo, hi = [int(v, 16) for v in rg.split("-")]
    sel = [r for r in recs if lo <= r[0] <= hi]
    ex = sum(r[3] for r in sel); sm = sum(r[2] for r in sel); fp = sum(r[3] for r in sel if is_fp64(r[1]))
    agg = {}
    for r in sel:
        for k, v in r[4].items(): agg[k] = agg.get(k, 0) + v
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{lo:#x}-{hi:#x}] {len(sel)} instrs, executed {ex/1e6:.1f} M ({100*ex/tot_exec:.1f} %), fp64 {fp/1e6:.1f} M, samples {sm} ({100*sm/max(1,tot_samp):.1f} %)  " + " ".join(f"{k[6:]}={v}" for k, v in top))
if a.top:
    for r in sorted(recs, key=lambda r: -r[2])[:a.top]:
        top = sorted(r[4].items(), key=lambda kv: -kv[1])[:3]
        print(f"{r[0]:#06x} samp {r[2]:6d} exec {r[3]/1e6:8.2f}M  {r[1][:70]:70s} " + " ".join(f"{k[6:]}={v}" for k, v in top))
if a.dump:
    lo, hi = [int(v, 16) for v in a.dump.split("-")]
    for r in recs:
        if lo <= r[0] <= hi:
            top = sorted(r[4].items(), key=lambda kv: -kv[1])[:2]
            print(f"{r[0]:#06x} samp {r[2]:5d} exec {r[3]/1e6:8.2f}M  {r[1][:80]:80s} " + " ".join(f"{k[6:]}={v}" for k, v in top if v))
