#!/usr/bin/env python
"""Quick A/B timing of kernel configurations on one GPU (not the contract bench).
usage: quick_bench.py [--size large] [--grid unionized] [--kernels 0,6] CONFIG...
  CONFIG = comma-separated ENV=VALUE pairs, e.g.  XSB200_GATHER=1,XSB200_BLOCKS_PER_SM=2
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xsbench_b200 as xs

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="large")
ap.add_argument("--grid", default="unionized")
ap.add_argument("--kernels", default="0")
ap.add_argument("--lookups", type=int, default=17_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--method", default="event")
ap.add_argument("configs", nargs="*", default=[""])
a = ap.parse_args()

extra = ["-l", str(a.lookups)] if a.method == "event" else []
inp = xs.read_CLI(["-s", a.size, "-m", a.method, "-G", a.grid] + extra)
sd = xs.grid_init_do_not_profile(inp)
expected = xs.expected_checksum(inp)
for cfg in a.configs:
    pairs = [kv.split("=") for kv in cfg.split(",") if "=" in kv]
    for k, v in pairs:
        os.environ[k] = v
    gpu = xs.move_simulation_data_to_device(inp, sd)
    for kid in [int(k) for k in a.kernels.split(",")]:
        i2 = xs.read_CLI(["-s", a.size, "-m", a.method, "-G", a.grid, "-k", str(kid)] + extra)
        best, res = 1e9, None
        for r in range(a.reps + 1):
            res = gpu.run(i2)
            if r:
                best = min(best, res.device_seconds)
        ok = "ok" if expected is None or res.checksum == expected else f"BAD({res.checksum})"
        ph = " ".join(f"{1e3*x:.2f}" for x in res.phase_seconds)
        print(f"{cfg or 'default':50s} k{kid} {1e3*best:8.3f} ms  {res.n_lookups/best/1e6:9.1f} M/s  phases[ms] {ph}  launches {res.gpu_launches} checksum {ok}", flush=True)
    gpu.release()
    for k, v in pairs:
        os.environ.pop(k, None)
