#!/usr/bin/env python
"""One xs_gpu_lookup_samples call on the non-fuel (or SEL=fuel / matK) samples of the canonical workload: an ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import xsbench_b200 as xs
os.environ["XSB200_E2E_CHUNKS"] = "1"
n = 17_000_000
inp = xs.read_CLI(["-s", "large", "-m", "event", "-G", "unionized", "-l", str(n)])
sd = xs.grid_init_do_not_profile(inp)
gpu = xs.move_simulation_data_to_device(inp, sd)
e, m, _, _ = gpu.dump(0, n)
which = os.environ.get("SEL", "non-fuel")
sel = (m != 0) if which == "non-fuel" else (m == 0) if which == "fuel" else (m == int(which[3:]))
ee, mm = np.ascontiguousarray(e[sel]), np.ascontiguousarray(m[sel])
for _ in range(int(os.environ.get("REPS", "2"))):
    r, _ = gpu.lookup_samples(ee, mm)
    print(which, len(ee), "lookup phase ms", 1e3 * r.phase_seconds[2], flush=True)
