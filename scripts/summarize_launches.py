#!/usr/bin/env python
"""Summarise an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`) of bench.py: per-kernel share of ONE timed step of the -k 4 pipeline.
usage: summarize_launches.py profiles/rNN_launches.csv rNN"""
import collections, csv, json, os, sys
path, tag = sys.argv[1], sys.argv[2]
out_dir = os.path.dirname(os.path.abspath(path))
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
launches = list(d.items())
names = [k[1] for k, _ in launches]
starts = [i for i, n in enumerate(names) if n.startswith('xs::xs_sample_kernel')]
i0 = starts[2]                                   # steps: warm-up, timed 1, timed 2 -> take timed 2
i1 = i0
while i1 < len(launches) and any(t in names[i1] for t in ('xs_sample_kernel', 'xs_partition_kernel', 'xs_window_kernel')) and (i1 == i0 or 'xs_sample_kernel' not in names[i1]):
    i1 += 1
st = launches[i0:i1]
agg = collections.OrderedDict()
for (i, n), m in st:
    short = n.split('(')[0].replace('void ', '')
    a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += m['gpu__time_duration.sum']; a[2] += m['dram__bytes_read.sum']; a[3] += m['dram__bytes_write.sum']
tot = sum(a[1] for a in agg.values())
L = [f"# {tag} launch list summary: one timed step of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` under ncu", "",
     f"Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/{tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (full list: `{os.path.basename(path)}`).",
     f"Times under ncu are cold-cache and serialised: compare shares, not absolutes (un-profiled numbers: `{tag}_bench.json`).", "",
     "| kernel | launches/step | time [us] | share | DRAM read [MB] | DRAM write [MB] |", "|---|---|---|---|---|---|"]
for k, a in agg.items():
    L.append(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f} % | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} |")
L.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot/1e3:.1f} | 100 % | {sum(a[2] for a in agg.values())/1e6:.1f} | {sum(a[3] for a in agg.values())/1e6:.1f} |")
w = [v for k, v in agg.items() if 'window' in k][0]
L += ["", f"Window kernel: {w[0]} launches per step (fuel windows + 1 launch for the other 11 materials), {100*w[1]/tot:.1f} % of the step; "
      f"DRAM traffic {(w[2]+w[3])/1e9:.2f} GB per step against 97.1 GB of algorithmic gather bytes (SURVEY 8d): the pair records are served "
      "from L2; DRAM streams the index rows, samples and partial sums.", "", "Per-launch detail of that step:", ""]
for (i, n), m in st:
    L.append(f"- #{i} `{n.split('(')[0].replace('void ', '')}`: {m['gpu__time_duration.sum']/1e3:.1f} us, DRAM read {m['dram__bytes_read.sum']/1e6:.0f} MB, write {m['dram__bytes_write.sum']/1e6:.0f} MB")
open(os.path.join(out_dir, f"{tag}_launches_summary.md"), 'w').write("\n".join(L) + "\n")
json.dump({"kernel": "xs_window_kernel<unionized>: all launches of one step (fuel windows + 1 launch for the other 11 materials)",
           "dram_bytes_per_launch": w[2] + w[3], "launches_per_step": w[0], "window_kernel_share_of_step": w[1] / tot,
           "what_bounds_it": "ncu --set full of one fuel window (profiles/%s_window_kernel_ncu.txt): l1tex__throughput 81 %%, "
                             "lts__throughput 70 %%, L2 hit 89 %%, issue slots 49 %% -- the L1 data path, not HBM" % tag,
           "source": f"profiles/{os.path.basename(path)} (ncu, one timed step of bench.py)",
           "dram_bytes_per_step_all_kernels": sum(a[2] + a[3] for a in agg.values())},
          open(os.path.join(out_dir, "lookup_kernel_traffic.json"), 'w'), indent=1)
print("\n".join(L[:14]))
