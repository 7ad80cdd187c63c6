#!/usr/bin/env python
"""Minimal-sector gather bytes of the reference layout: how many distinct 32-byte
sectors ONE lookup must touch when nothing is shared between lookups -- the UEG binary-search
probes, the index-row entries of the material's nuclides and the low/high grid points of every
nuclide (2 x 48 B, 16-byte aligned).  Host-only: uses the host generator (libxsb200_host.so) and
numpy; the sample stream is the reference's (LCG, seed 1070).

usage: sector_counter.py [--size small|large] [--lookups 200000]
Prints bytes/lookup (algorithmic and in sectors) and the lookups/s a perfect HBM gather of those
sectors would reach at the measured copy bandwidth (MEASURED_PEAKS.json, else 6539.9 GB/s).
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xsbench_b200 as xs
from xsbench_b200 import _abi
import ctypes as C

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="large")
ap.add_argument("--lookups", type=int, default=200_000)
a = ap.parse_args()

inp = xs.read_CLI(["-s", a.size, "-m", "event", "-G", "unionized"])
sd = xs.grid_init_do_not_profile(inp)
arr = xs.simulation_arrays(inp, sd)
n_iso, n_gp = inp.n_isotopes, inp.n_gridpoints
ueg, index = arr["unionized_energy_array"], arr["index_grid"].reshape(-1, n_iso)
num_nucs, mats = arr["num_nucs"], arr["mats"].reshape(12, -1)
n_ueg = len(ueg)

# the reference's sample stream (cuda/Simulation.cu:53-60): seed 1070 advanced 2*i, energy then material
lib = _abi.host_lib()
lib.fast_forward_LCG.restype = C.c_uint64
lib.fast_forward_LCG.argtypes = [C.c_uint64, C.c_uint64]
lib.LCG_random_double.restype = C.c_double
lib.pick_mat.restype = C.c_int
e = np.empty(a.lookups); m = np.empty(a.lookups, np.int64)
seed = C.c_uint64(1070)
for i in range(a.lookups):          # sequential draws == fast_forward(2*i) per lookup
    e[i] = lib.LCG_random_double(C.byref(seed))
    m[i] = lib.pick_mat(C.byref(seed))

SECTOR = 32
total_sectors = np.zeros(a.lookups, np.int64)

# 1. UEG binary search (cuda/Simulation.cu:241-261): probe addresses mid*8
lo = np.zeros(a.lookups, np.int64); hi = np.full(a.lookups, n_ueg - 1, np.int64)
probes = []
while np.any(hi - lo > 1):
    active = hi - lo > 1
    mid = lo + (hi - lo) // 2
    probes.append(np.where(active, mid * 8 // SECTOR, -1))
    up = ueg[mid] > e
    hi = np.where(active & up, mid, hi)
    lo = np.where(active & ~up, mid, lo)
probes = np.stack(probes, 1)
probes.sort(1)
search_sectors = ((probes[:, 1:] != probes[:, :-1]) & (probes[:, 1:] >= 0)).sum(1) + (probes[:, 0] >= 0)
total_sectors += search_sectors
row = lo
search_levels = probes.shape[1]

# 2 + 3. per material: index entries of the row, low/high grid points of every nuclide
alg_bytes = np.zeros(a.lookups)
for mat in range(12):
    sel = np.nonzero(m == mat)[0]
    if len(sel) == 0:
        continue
    nucs = mats[mat, :num_nucs[mat]].astype(np.int64)
    r = row[sel][:, None]
    idx_sector = (r * n_iso + nucs[None, :]) * 4 // SECTOR
    idx_sector.sort(1)
    total_sectors[sel] += (idx_sector[:, 1:] != idx_sector[:, :-1]).sum(1) + 1
    low = index[row[sel]][:, nucs].astype(np.int64)
    low = np.where(low == n_gp - 1, n_gp - 2, low)
    first_byte = (nucs[None, :] * n_gp + low) * 48
    s0, s1 = first_byte // SECTOR, (first_byte + 95) // SECTOR
    total_sectors[sel] += (s1 - s0 + 1).sum(1)          # distinct nuclides never share a sector pair
    alg_bytes[sel] = len(nucs) * (4 + 96) + search_levels * 8

peak = 6539.9
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
bytes_sector = total_sectors.mean() * SECTOR
print(f"{a.size}: {a.lookups} lookups, mean nuclides per lookup {np.mean([num_nucs[k] for k in m]):.4f}")
print(f"algorithmic gather bytes per lookup  : {alg_bytes.mean():9.1f} B   (N_nuc*(4+2*48) + {search_levels} search levels * 8)")
print(f"minimal distinct 32-B sectors        : {total_sectors.mean():9.2f}   = {bytes_sector:9.1f} B per lookup (x{bytes_sector/alg_bytes.mean():.3f})")
print(f"  of which UEG search probes         : {search_sectors.mean():9.2f}   (the top levels are shared by all lookups and live in cache; without them: "
      f"{(total_sectors - search_sectors).mean() * SECTOR:.1f} B per lookup)")
print(f"perfect HBM gather at {peak:.1f} GB/s   : {peak*1e9/alg_bytes.mean()/1e9:6.3f} G lookups/s algorithmic, {peak*1e9/bytes_sector/1e9:6.3f} G lookups/s in sectors")
xs.free_simulation_data(sd)
