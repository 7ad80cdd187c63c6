#!/usr/bin/env python
"""A/B timing of the host-buffer path (xs_gpu_lookup_samples) for env configurations.
usage: e2e_bench.py CONFIG...   (CONFIG = comma-separated ENV=VALUE pairs)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xsbench_b200 as xs

n = 17_000_000
inp = xs.read_CLI(["-s", "large", "-m", "event", "-G", "unionized", "-l", str(n)])
sd = xs.grid_init_do_not_profile(inp)
e_pin = m_pin = None
for cfg in (sys.argv[1:] or [""]):
    pairs = [kv.split("=") for kv in cfg.split(",") if "=" in kv]
    for k, v in pairs:
        os.environ[k] = v
    gpu = xs.move_simulation_data_to_device(inp, sd)
    if e_pin is None:
        e, m, _, _ = gpu.dump(0, n)
        e_pin, m_pin = torch.from_numpy(e).pin_memory(), torch.from_numpy(m).pin_memory()
    for _ in range(2):
        gpu.lookup_samples(None, None, n=n, energy_ptr=e_pin.data_ptr(), mat_ptr=m_pin.data_ptr())
    t0 = time.perf_counter()
    for _ in range(5):
        r, _ = gpu.lookup_samples(None, None, n=n, energy_ptr=e_pin.data_ptr(), mat_ptr=m_pin.data_ptr())
    dt = (time.perf_counter() - t0) / 5
    print(f"{cfg or 'default':40s} {1e3*dt:7.3f} ms/step  {n/dt/1e6:8.1f} M lookups/s  device {1e3*r.device_seconds:.3f} ms  checksum {r.checksum}", flush=True)
    gpu.release()
    for k, v in pairs:
        os.environ.pop(k, None)
