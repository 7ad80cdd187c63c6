#!/usr/bin/env python
"""Summarise an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`) of bench.py: per-kernel share of ONE timed step of the headline
pipeline (-k 6: sample -> radix sort -> lane-per-lookup kernels xs_dense_kernel + xs_sorted_kernel).
usage: summarize_launches.py profiles/rNN_launches.csv rNN [bound-summary text]"""
import collections, csv, json, os, sys
path, tag = sys.argv[1], sys.argv[2]
bound = sys.argv[3] if len(sys.argv) > 3 else None
out_dir = os.path.dirname(os.path.abspath(path))
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(',', ''))
launches = list(d.items())
names = [k[1] for k, _ in launches]
STEP = ('xs_sample_kernel', 'sort_', 'onesweep_', 'xs_build_segments_kernel', 'xs_gather_kernel', 'xs_bin_scatter', 'xs_partition_kernel', 'xs_window_kernel', 'xs_sorted_kernel', 'xs_dense_kernel')
starts = [i for i, n in enumerate(names) if 'xs_sample_kernel' in n]
i0 = starts[2]                                   # steps: warm-up, timed 1, timed 2 -> take timed 2
i1 = i0 + 1
while i1 < len(launches) and any(t in names[i1] for t in STEP) and 'xs_sample_kernel' not in names[i1]:
    i1 += 1
st = launches[i0:i1]
agg = collections.OrderedDict()
for (i, n), m in st:
    short = n.split('(')[0].replace('void ', '').replace('xs::', '')
    a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += m['gpu__time_duration.sum']; a[2] += m['dram__bytes_read.sum']; a[3] += m['dram__bytes_write.sum']
tot = sum(a[1] for a in agg.values())
L = [f"# {tag} launch list summary: one timed step of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-traffic-probe --no-extras` under ncu", "",
     f"Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/{tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-traffic-probe --no-extras` (full list: `{os.path.basename(path)}`).",
     f"Times under ncu are cold-cache and serialised: compare shares, not absolutes (un-profiled numbers: `{tag}_bench.json`).", "",
     "| kernel | launches/step | time [us] | share | DRAM read [MB] | DRAM write [MB] |", "|---|---|---|---|---|---|"]
for k, a in agg.items():
    L.append(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f} % | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} |")
L.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot/1e3:.1f} | 100 % | {sum(a[2] for a in agg.values())/1e6:.1f} | {sum(a[3] for a in agg.values())/1e6:.1f} |")
dom_name, w = max(agg.items(), key=lambda kv: kv[1][1])
L += ["", f"Dominant kernel `{dom_name}`: {w[0]} launch(es) per step, {100*w[1]/tot:.1f} % of the step; DRAM traffic "
      f"{(w[2]+w[3])/1e9:.2f} GB per launch against 97.1 GB of algorithmic gather bytes (SURVEY 8d): the pair records are shared by "
      "the lookups of a warp and served from shared memory / L2; DRAM carries the records' first touch, the (randomly placed) samples and two index-row segments per warp-group and chunk.",
      "", "Per-launch detail of that step:", ""]
for (i, n), m in st:
    L.append(f"- #{i} `{n.split('(')[0].replace('void ', '')}`: {m['gpu__time_duration.sum']/1e3:.1f} us, DRAM read {m['dram__bytes_read.sum']/1e6:.0f} MB, write {m['dram__bytes_write.sum']/1e6:.0f} MB")
open(os.path.join(out_dir, f"{tag}_launches_summary.md"), 'w').write("\n".join(L) + "\n")
json.dump({"kernel": f"{dom_name}: {w[0]} launch(es) per step",
           "dram_bytes_per_launch": (w[2] + w[3]) / w[0], "launches_per_step": w[0], "share_of_step": w[1] / tot,
           "what_bounds_it": bound or "see profiles/%s_notes.md" % tag,
           "source": f"profiles/{os.path.basename(path)} (ncu, one timed step of bench.py)",
           "lookup_phase": "xs_dense_kernel + xs_sorted_kernel (2 launches per step)",
           "lookup_phase_dram_bytes": sum(a[2] + a[3] for k, a in agg.items() if 'xs_dense_kernel' in k or 'xs_sorted_kernel' in k or 'xs_window_kernel' in k),
           "lookup_phase_share_of_step": sum(a[1] for k, a in agg.items() if 'xs_dense_kernel' in k or 'xs_sorted_kernel' in k or 'xs_window_kernel' in k) / tot,
           "dram_bytes_per_step_all_kernels": sum(a[2] + a[3] for a in agg.values())},
          open(os.path.join(out_dir, "lookup_kernel_traffic.json"), 'w'), indent=1)
print("\n".join(L[:20]))
