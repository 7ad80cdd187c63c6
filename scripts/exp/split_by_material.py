#!/usr/bin/env python
"""Lookup-phase time of the -k 6 pipeline for fuel-only / non-fuel-only / all samples (host-sample API)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import xsbench_b200 as xs
os.environ["XSB200_E2E_CHUNKS"] = "1"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 17_000_000
per_material = len(sys.argv) <= 2
inp = xs.read_CLI(["-s", "large", "-m", "event", "-G", "unionized", "-l", str(n)])
sd = xs.grid_init_do_not_profile(inp)
gpu = xs.move_simulation_data_to_device(inp, sd)
e, m, _, _ = gpu.dump(0, n)
for name, sel in (("all", slice(None)), ("fuel", m == 0), ("non-fuel", m != 0)) + (tuple((f"mat{k}", m == k) for k in range(1, 12)) if per_material else ()):
    ee, mm = np.ascontiguousarray(e[sel]), np.ascontiguousarray(m[sel])
    best = None
    for _ in range(3):
        r, _ = gpu.lookup_samples(ee, mm)
        ph = r.phase_seconds
        best = ph if best is None or ph[2] < best[2] else best
    print(f"{name:9s} {len(ee):9d} lookups: locate {1e3*best[0]:.2f}  sort {1e3*best[1]:.2f}  lookup {1e3*best[2]:.3f} ms", flush=True)
