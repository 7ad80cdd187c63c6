#!/usr/bin/env python
"""bench.py -- XS lookups/s on the canonical XSBench configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): -s large -m event -G unionized, 17,000,000 lookups
(355 nuclides x 11,303 grid points, 5.6 GB of cross-section data, synthetic, generated exactly
as the reference does).  One "step" = one pass of the lookup path over 17 M lookups per GPU.
Multi-GPU is weak scaling: rank r performs lookup ids [r*17M, (r+1)*17M) on its own replica of the
grid (xsbench_b200/sharding.py); the only collective is the all-reduce of {verification sum,
lookup count}, inside the timed region.

value : device-timed lookups/s of the whole pipeline [sample -> sort by (material, energy) ->
        lane-per-lookup kernels -> checksum] (-k 6 semantics: like the reference's own fastest
        variant, the sort is inside the timed region); the grid is resident, nothing else is.
e2e   : the same lookups through the host-buffer C-ABI call xs_gpu_lookup_samples: the caller's
        (f64 energy, int material) samples in pinned host memory -- 12 B/lookup -- are copied each
        step (9 B/lookup cross PCIe: the library narrows the materials to bytes on the host while
        the energies are in flight), the result is read back each step.
roofline: what BINDS the dominant kernel.  For -k 6 that is FP64 issue, not HBM: the sorted
        lookups of a warp share their grid records, so DRAM moves ~3 % of the algorithmic gather
        bytes; the floor is (lookup, nuclide) pairs x FP64 operations per pair / (SMs x 64 FP64
        lanes/clk x SM clock).  `frac` = that floor / the measured lookup-phase time.  The HBM view
        (algorithmic bytes of SURVEY.md 8(d), measured DRAM traffic) is reported next to it, and
        `variants.k0` -- which IS a DRAM-bound gather -- carries its own HBM roofline.
gpu_baseline: the UNMODIFIED reference cuda/ build (oracle/_ref/XSBench_cuda_ref, sm_100a) run as
        a subprocess on the same GPU in the same run, -k 0 and -k 6 (its own host timer), next to
        its device-timed kernel sums from the committed ncu launch list.
cpu_baseline / --impl reference: the UNMODIFIED reference openmp-threading code compiled into
        oracle/_ref/libxsref.so, on the same generated data, all host cores: its sorted variant
        run_event_based_simulation_optimization_1 (its "-k 1": material + energy sort -- what
        cuda "-k 6" is) when --kernel 6, else run_event_based_simulation; falls back to the oracle
        port when that library did not travel.
strong / bands (extras of the same line): 10^9 lookups of the same problem split N ways (fixed total:
        strong scaling, efficiency against this run's own 1-GPU time), and -- N >= 2 -- the unionized
        index grid sharded by energy band, one band per rank (what XXL needs), checksum all-reduced.
"""
from __future__ import annotations

import argparse
import ctypes as C
import csv
import json
import math
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOOKUPS_PER_GPU = 17_000_000
STRONG_LOOKUPS = 1_000_000_000
FP64_LANES_PER_CLK_PER_SM = 64          # B200: DADD / DMUL / DFMA all issue 2 warp-instructions/clk/SM (scripts/exp/fp64_ops.cu)
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "XSBench_cuda_ref")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="large", help="problem size (large is the headline; small for quick checks)")
    ap.add_argument("--kernel", type=int, default=6,
                    help="-k variant timed as `value`: 6 = (material, energy) sort + lane-per-lookup kernels (fastest, default); "
                         "4 = material grouping + windowed nuclide sweep; 0 = fused in-order kernel (baseline semantics)")
    ap.add_argument("--lookups", type=int, default=LOOKUPS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference cuda/ subprocess runs")
    ap.add_argument("--no-traffic-probe", action="store_true", help="skip the ncu DRAM-traffic pass (falls back to profiles/)")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling / energy-band extras")
    ap.add_argument("--strong-lookups", type=int, default=STRONG_LOOKUPS)
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)   # internal: the process ncu profiles
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks (recipe in B200_PROFILING.md): sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region.  Uses NVML from a
    background thread (measured: polling `nvidia-smi -lms 20` slowed the timed kernels by 8 %;
    NVML queries every 25 ms do not), and falls back to the nvidia-smi recipe line."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []          # (time, sm_mhz, sm_max_mhz, power_w, reasons bitmask or list)
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            index = self.gpu_index
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                try:
                    index = int(visible.split(",")[self.gpu_index])
                except (ValueError, IndexError):
                    pass
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"
            self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        names = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.time(), float(sm), float(self.sm_max), pw, [k for k, bit in names.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.025)

    def _pump_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = line.strip().split(", ")
            if len(r) < 9:
                continue
            try:
                self.samples.append((time.time(), float(r[1]), float(r[2]), float(r[3]),
                                     [k for k, v in zip(names, r[5:9]) if v.strip().lower() == "active"]))
            except ValueError:
                continue

    def stop(self, t0: float, t1: float):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.06)
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for r in self.samples if t0 <= r[0] <= t1 + 0.05] or self.samples
        reasons = sorted({k for r in rows for k in r[4]})
        return {"sm_mhz": statistics.median(r[1] for r in rows) if rows else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "power_w_max": max(r[3] for r in rows) if rows else None,
                "samples": len(rows), "source": self.mode, "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(inp, sd, n_lookups: int, kernel: int):
    """Time the reference's own CPU driver on `sd`: its sorted variant (openmp-threading "-k 1",
    Simulation.c:698-871) when `kernel` is 6, else run_event_based_simulation (:15-114).
    Returns (lookups/s, kind, cores, checksum, seconds, variant)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    if ol.have_reference():
        r = ol.reference()
        rinp = ol.RefInputs(cores, inp.n_isotopes, inp.n_gridpoints, n_lookups, inp.HM, inp.grid_type,
                            inp.hash_bins, 0, 2, 0, 1 if kernel == 6 else 0)
        rsd = ol.RefSimulationData()
        for name in ("num_nucs", "concs", "mats", "unionized_energy_array", "index_grid"):
            setattr(rsd, name, getattr(sd, name))
        rsd.nuclide_grid = C.cast(sd.nuclide_grid, C.POINTER(C.c_double))
        for name in ("length_num_nucs", "length_concs", "length_mats", "length_unionized_energy_array",
                     "length_index_grid", "length_nuclide_grid", "max_num_nucs"):
            setattr(rsd, name, getattr(sd, name))
        fn = r.run_event_based_simulation_optimization_1 if kernel == 6 else r.run_event_based_simulation
        variant = ("openmp-threading run_event_based_simulation_optimization_1 (-k 1: material + energy sort, the CPU "
                   "counterpart of cuda -k 6)" if kernel == 6 else "openmp-threading run_event_based_simulation (-k 0)")
        t0 = time.perf_counter()
        v = fn(rinp, rsd, 1)                                 # mype=1: no progress print
        dt = time.perf_counter() - t0
        return n_lookups / dt, "reference", cores, int(v), dt, variant
    # the compiled reference did not travel: time the oracle port on its own generated data
    p = ol.OracleProblem(inp.n_isotopes, inp.n_gridpoints, inp.grid_type, inp.hash_bins)
    t0 = time.perf_counter()
    v = p.event(0, n_lookups, cores)
    dt = time.perf_counter() - t0
    p.close()
    return n_lookups / dt, "port", cores, int(v), dt, "oracle port of run_event_based_simulation (-k 0)"


# ------------------------------------------------------------------------------------------------
# reference cuda/ build on the same GPU (reported baseline)
# ------------------------------------------------------------------------------------------------
def gpu_reference_baseline(size: str, lookups: int, device_index: int):
    """Run the unmodified reference cuda/ binary (compiled for sm_100a by oracle/Makefile) for -k 0 and
    -k 6 and parse its own report (host-timed: cuda/Main.cu:59-97 brackets kernels + sync + Thrust
    reduce, and for -k 6 the cudaMallocs and sorts).  Device-timed kernel sums come from the committed
    ncu launch list of the same binary (profiles/r02_reference_cuda_device_times.json)."""
    if not os.path.exists(REF_CUDA):
        return {"unavailable": "oracle/_ref/XSBench_cuda_ref did not travel (built by `make -C oracle ref_cuda` where /root/reference exists)"}
    out = {"binary": "oracle/_ref/XSBench_cuda_ref (unmodified reference cuda/, -gencode arch=compute_100a,code=sm_100a)",
           "timer": "its own host timer (cuda/Main.cu:59-97)"}
    env = dict(os.environ)
    visible = env.get("CUDA_VISIBLE_DEVICES")
    if visible:
        try:
            env["CUDA_VISIBLE_DEVICES"] = visible.split(",")[device_index]
        except IndexError:
            pass
    else:
        env["CUDA_VISIBLE_DEVICES"] = str(device_index)
    for k in (0, 6):
        best = None
        for _ in range(2):      # its timer includes cudaMalloc / Thrust temporaries (a cold first process has read 0.13 s for -k 6): best of two
            try:
                p = subprocess.run([REF_CUDA, "-m", "event", "-s", size, "-l", str(lookups), "-k", str(k)], env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
                rate = re.search(r"Lookups/s:\s+([\d,]+)", p.stdout)
                chk = re.search(r"Verification checksum:\s+(\d+)\s+\((\w+)\)", p.stdout)
                rt = re.search(r"Runtime:\s+([\d.]+) seconds", p.stdout)
                one = {"lookups_per_s": float(rate.group(1).replace(",", "")) if rate else None,
                       "runtime_s": float(rt.group(1)) if rt else None,
                       "checksum": int(chk.group(1)) if chk else None, "valid": bool(chk and chk.group(2) == "Valid"),
                       "runs": "best of 2 processes"}
                if best is None or (one["lookups_per_s"] or 0) > (best.get("lookups_per_s") or 0):
                    best = one
            except Exception as exc:                         # noqa: BLE001 -- a baseline must never take the bench down
                best = best or {"error": repr(exc)[:200]}
        out[f"k{k}"] = best
    prof = os.path.join(ROOT, "profiles", "r02_reference_cuda_device_times.json")
    if os.path.exists(prof):
        try:
            out["device_timed"] = json.load(open(prof))
        except Exception:
            pass
    return out


# ------------------------------------------------------------------------------------------------
# DRAM traffic of the lookup kernels: one ncu pass over a short run of this very file
# ------------------------------------------------------------------------------------------------
def traffic_probe_main(args):
    """The process ncu profiles (bench.py --traffic-probe): one warm-up and one measured pass of -k 6,
    -k 0 and -k 4 on the device-generated problem.  Prints nothing the parent parses; ncu's csv is the output."""
    import xsbench_b200 as xs
    inp = xs.read_CLI(["-s", args.size, "-m", "event", "-G", "unionized", "-l", str(args.lookups), "-k", "6"])
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        for k in (6, 6, 0, 0, 4):
            gpu.run(xs.read_CLI(["-s", args.size, "-m", "event", "-G", "unionized", "-l", str(args.lookups), "-k", str(k)]))
    xs.free_simulation_data(mats)
    return 0


def measure_traffic(args, device_index: int):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the last launch of every lookup
    kernel, from `ncu` run on a subprocess of this file.  Returns None when ncu is not usable."""
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if not ncu:
        return None
    tmp = tempfile.mkdtemp(prefix="xsb200_traffic_")
    log = os.path.join(tmp, "launches.csv")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--kernel-name", "regex:xs_(dense|sorted|event|window|tile)_kernel", "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--traffic-probe", "--size", args.size, "--lookups", str(args.lookups)]
    try:
        p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=420)
        if p.returncode != 0 or not os.path.exists(log):
            return None
        rows = [r for r in csv.reader(open(log, errors="replace")) if len(r) > 5]
        hdr = rows[0]
        ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
        per = {}
        for r in rows[1:]:
            try:
                per.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
            except ValueError:
                continue
        launches = sorted(per.items())
        def last_of(tag, how_many=1):
            hits = [m for (i, n), m in launches if tag in n]
            hits = hits[-how_many:] if hits else []
            return (sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in hits) if hits else None,
                    sum(m.get("gpu__time_duration.sum", 0.0) for m in hits) * 1e-9 if hits else None)
        n_window = sum(1 for (i, n), m in launches if "xs_window_kernel" in n)
        dense_b, dense_t = last_of("xs_dense_kernel")
        sparse_b, sparse_t = last_of("xs_sorted_kernel")
        event_b, event_t = last_of("xs_tile_kernel")
        if event_b is None:
            event_b, event_t = last_of("xs_event_kernel")
        window_b, window_t = last_of("xs_window_kernel", n_window)
        return {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on a subprocess of this bench (one pass each of -k 6 / -k 0 / -k 4, same problem), this run",
                "k6_lookup_phase_bytes": (dense_b or 0.0) + (sparse_b or 0.0) if dense_b is not None or sparse_b is not None else None,
                "k6_dense_kernel_bytes": dense_b, "k6_dense_kernel_s_under_ncu": dense_t,
                "k6_sparse_kernel_bytes": sparse_b, "k0_kernel_bytes": event_b, "k0_kernel_s_under_ncu": event_t,
                "k4_lookup_phase_bytes": window_b}
    except Exception:
        return None
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def committed_traffic():
    prof = os.path.join(ROOT, "profiles", "lookup_kernel_traffic.json")
    if not os.path.exists(prof):
        return None
    try:
        pj = json.load(open(prof))
        return {"source": "profiles/lookup_kernel_traffic.json (committed ncu capture of the same command; the live probe did not run)",
                "k6_lookup_phase_bytes": pj.get("lookup_phase_dram_bytes", pj.get("dram_bytes_per_launch")),
                "k0_kernel_bytes": pj.get("k0_kernel_dram_bytes")}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.traffic_probe:
        return traffic_probe_main(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and world != args.gpus:
        args.gpus = world

    import xsbench_b200 as xs
    from xsbench_b200 import sharding

    workload = f"-s {args.size} -m event -G unionized -l {args.lookups} -k {args.kernel}"

    # ---------------- reference arm: CPU, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        inp = xs.make_inputs(size=args.size, method="event", grid="unionized", lookups=args.lookups)
        sd = xs.grid_init_do_not_profile(inp)
        # size the per-step sample so the whole run stays within a few minutes
        probe_n = min(args.lookups, 500_000)
        rate, kind, cores, _, _, variant = cpu_reference_rate(inp, sd, probe_n, args.kernel)
        budget_s = 150.0 / max(1, args.steps + args.warmup)
        sample = int(min(args.lookups, max(200_000, rate * min(budget_s, 20.0))))
        times, checksum = [], None
        for s in range(args.warmup + args.steps):
            r, kind, cores, v, dt, variant = cpu_reference_rate(inp, sd, sample, args.kernel)
            if s >= args.warmup:
                times.append(dt)
            checksum = v % 999983
        value = sample * len(times) / sum(times)
        line = {
            "impl": "reference", "metric": "XS lookups/sec", "value": value, "unit": "lookups/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": f"first {sample} lookups per step", "threads": cores,
                       "reference_variant": variant},
            "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": cores, "kind": kind, "variant": variant,
                             "sample": f"first {sample} of {args.lookups} lookups of the same workload per step"},
            "e2e": {"value": value, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "checksum_of_sample": checksum,
        }
        print(json.dumps(line))
        return 0

    # ---------------- our arm ----------------
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the lookup path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    ncpu = os.cpu_count() or 1
    gen_threads = max(1, ncpu // max(1, world))

    def cli(kernel, lookups=args.lookups):
        return xs.read_CLI(["-s", args.size, "-m", "event", "-G", "unionized", "-l", str(lookups), "-k", str(kernel),
                            "-t", str(gen_threads)])

    inp = cli(args.kernel)
    t_gen = time.perf_counter()
    # N = 1 keeps the host copy of the problem for the CPU baseline; with several ranks every
    # rank builds its replica directly on its GPU (byte-identical device-side generator)
    sd = xs.grid_init_do_not_profile(inp) if world == 1 else xs.materials_only(inp)
    t_gen = time.perf_counter() - t_gen
    num_nucs = xs.simulation_arrays(inp, sd)["num_nucs"].astype(np.int64).copy()
    t_up = time.perf_counter()
    gpu = xs.move_simulation_data_to_device(inp, sd)
    t_up = time.perf_counter() - t_up
    stream = torch.cuda.current_stream()
    gpu.set_stream(stream.cuda_stream)
    info = gpu.info()

    first_id, n_mine = sharding.weak_shard(args.lookups, rank, world)   # weak scaling: every rank owns 17 M distinct lookups
    total_lookups = args.lookups * world
    reducer = sharding.ResultReducer("cuda")

    def step():
        res = gpu.run_range(first_id, n_mine)
        reducer.submit(res.verification, res.n_lookups)     # the path's only collective (NCCL all-reduce, 16 bytes)
        return res

    for _ in range(args.warmup):
        step()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record(stream)
    lookup_kernel_s, phase_s, launches = [], [], 0
    for _ in range(args.steps):
        res = step()
        lookup_kernel_s.append(res.phase_seconds[2])
        phase_s.append(res.phase_seconds)
        launches += res.gpu_launches
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    elapsed_s = max_over_ranks([ev0.elapsed_time(ev1) * 1e-3])[0]
    verification, counted = reducer.result()
    clock_info = clocks.stop(wall0, wall1) if rank == 0 else None

    def timed_runs(fn, n_runs):
        """Device time of n_runs calls of fn (after one un-timed call), max over ranks."""
        out = fn()
        barrier()
        ev0.record(stream)
        for _ in range(n_runs):
            out = fn()
        ev1.record(stream)
        barrier()
        return max_over_ranks([ev0.elapsed_time(ev1) * 1e-3])[0], out

    # ---------------- other kernel variants (informational, same timing protocol, fewer steps) -----
    variants = {}
    n_v = max(1, min(args.steps, 3))
    for kid in (0, 4, 6):
        if kid == args.kernel:
            continue
        vin = cli(kid)
        tv, rv = timed_runs(lambda: gpu.run_range(first_id, n_mine, vin), n_v)
        variants[f"k{kid}"] = {"lookups_per_s": total_lookups * n_v / tv, "ms_per_step": 1e3 * tv / n_v,
                               "lookup_phase_ms": 1e3 * rv.phase_seconds[2],
                               "checksum_matches": bool(rv.verification == res.verification)}

    # ---------------- e2e: host buffers through the C ABI ----------------
    # samples come from the device sampler once (synthetic), then live in pinned host memory
    e_host, m_host, _, _ = gpu.dump(first_id, n_mine)
    mat_hist = np.bincount(m_host, minlength=12).astype(np.int64)
    pairs_mine = int(np.dot(mat_hist, num_nucs))            # (lookup, nuclide) pairs of this rank's lookups: exact
    e_pin = torch.from_numpy(e_host).pin_memory()
    m_pin = torch.from_numpy(m_host).pin_memory()
    del e_host, m_host
    for _ in range(max(1, min(args.warmup, 2))):
        gpu.lookup_samples(None, None, n=n_mine, energy_ptr=e_pin.data_ptr(), mat_ptr=m_pin.data_ptr())
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_verification = None
    e2e_device_s = []
    for _ in range(e2e_steps):
        r2, _ = gpu.lookup_samples(None, None, n=n_mine, energy_ptr=e_pin.data_ptr(), mat_ptr=m_pin.data_ptr())
        reducer.submit(r2.verification, r2.n_lookups)
        h2d, d2h = r2.h2d_bytes, r2.d2h_bytes
        e2e_verification = r2.verification
        e2e_device_s.append(r2.device_seconds)
    barrier()
    e2e_s = max_over_ranks([time.perf_counter() - t0])[0]    # host wall clock: includes the copies and the sync
    e2e_ok = (e2e_verification == res.verification)
    # the copy alone, all ranks at once: what PCIe / the host's memory system allow this many GPUs of the node to
    # pull at the same time (the ceiling of the e2e number: at N = 8 the ranks share the host's root complexes)
    copy_floor = None
    try:
        d_e = torch.empty(n_mine, dtype=torch.float64, device="cuda")
        d_m = torch.empty(n_mine, dtype=torch.int32, device="cuda")
        best = None
        for _ in range(3):
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            d_e.copy_(e_pin, non_blocking=True)
            d_m.copy_(m_pin, non_blocking=True)
            c1.record()
            c1.synchronize()
            ms = max_over_ranks([c0.elapsed_time(c1)])[0]
            best = ms if best is None else min(best, ms)
        copy_floor = {"ms_12B_per_lookup": best, "gbs_per_gpu": 12 * n_mine / best / 1e6,
                      "what": "cudaMemcpyAsync of the same pinned (f64 energy, int material) arrays on all ranks at once, nothing else running: best of 3, max over ranks"}
        del d_e, d_m
    except Exception as exc:                                  # informational only
        copy_floor = {"error": str(exc)[:200]}
    del e_pin, m_pin

    # ---------------- the fused-arithmetic build of the dense kernel (informational) ----------------
    # XSB200_ARITH=fused is read by xs_gpu_init: a second context (device-generated replica of the same problem)
    if args.kernel == 6 and not args.no_extras:
        os.environ["XSB200_ARITH"] = "fused"
        try:
            fmats = xs.materials_only(inp)
            fgpu = xs.move_simulation_data_to_device(cli(6), fmats)
            fgpu.set_stream(stream.cuda_stream)
            fin = cli(6)
            tf, rf = timed_runs(lambda: fgpu.run_range(first_id, n_mine, fin), n_v)
            f_ops = getattr(fgpu.info(), "fp64_ops_per_pair", 12) or 12
            variants["k6_fused"] = {"lookups_per_s": total_lookups * n_v / tf, "ms_per_step": 1e3 * tf / n_v,
                                    "lookup_phase_ms": 1e3 * rf.phase_seconds[2], "fp64_ops_per_pair": f_ops,
                                    "checksum_matches": bool(rf.verification == res.verification),
                                    "note": "XSB200_ARITH=fused: 12 FP64 operations per (lookup, nuclide) instead of the reference's 24 roundings; "
                                            "macro_xs within a few ulp (contract 1e-12), integers guarded; not the default"}
            fgpu.release()
            xs.free_simulation_data(fmats)
        finally:
            os.environ.pop("XSB200_ARITH", None)

    # ---------------- extras: strong scaling of a fixed 10^9 lookups; energy-band sharding ----------------
    strong = bands = None
    if not args.no_extras and args.size == "large":
        total = args.strong_lookups
        sin = cli(6, lookups=total)
        s_first, s_count = sharding.strong_shard(total, rank, world)
        s_reducer = sharding.ResultReducer("cuda")

        def strong_run():
            r = gpu.run_range(s_first, s_count, sin)
            s_reducer.submit(r.verification, r.n_lookups)
            return r
        ts, _ = timed_runs(strong_run, 2)
        sv, sn = s_reducer.result()
        expected_strong = xs.expected_checksum(sin)
        strong = {"lookups": total, "scaling": "strong", "kernel": 6, "n_gpus": world, "seconds": ts / 2,
                  "value": total * 2 / ts, "unit": "lookups/s", "checksum": sv % 999983,
                  "checksum_expected": expected_strong, "checksum_ok": bool(expected_strong is not None and sv % 999983 == expected_strong),
                  "lookups_counted": sn}
        if world > 1:
            # this run's own 1-GPU time for the same total (rank 0 alone; the others wait at the barrier)
            t1 = 0.0
            if rank == 0:
                gpu.run_range(0, total, sin)
                torch.cuda.synchronize()
                ev0.record(stream)
                gpu.run_range(0, total, sin)
                ev1.record(stream)
                torch.cuda.synchronize()
                t1 = ev0.elapsed_time(ev1) * 1e-3
            barrier()
            t1 = max_over_ranks([t1])[0]
            strong["one_gpu_seconds"] = t1
            strong["speedup"] = t1 / (ts / 2)
            strong["efficiency"] = t1 / (ts / 2) / world

            # energy bands: rank r holds the index rows of band r only (XSB200_BANDS / XSB200_BAND_INDEX), draws
            # every one of the 17 M ids and keeps those whose row is in its band; the all-reduce adds the bands up
            os.environ["XSB200_BANDS"], os.environ["XSB200_BAND_INDEX"] = str(world), str(rank)
            try:
                bmats = xs.materials_only(inp)
                bgpu = xs.move_simulation_data_to_device(cli(6), bmats)
                bgpu.set_stream(stream.cuda_stream)
                b_reducer = sharding.ResultReducer("cuda")
                bin_ = cli(6)

                def band_run():
                    r = bgpu.run(bin_)
                    b_reducer.submit(r.verification, r.n_lookups)
                    return r
                tb, rb = timed_runs(band_run, 2)
                bv, bn = b_reducer.result()
                b_expected = xs.expected_checksum(bin_)
                bands = {"n_bands": world, "lookups": args.lookups, "seconds": tb / 2, "value": args.lookups * 2 / tb,
                         "unit": "lookups/s", "index_rows_bytes_per_gpu": int(bgpu.info().resident_bytes),
                         "lookups_kept_by_rank0": int(rb.n_lookups) if rank == 0 else None,
                         "checksum": bv % 999983, "checksum_expected": b_expected,
                         "checksum_ok": bool(bn == args.lookups and b_expected is not None and bv % 999983 == b_expected)}
                bgpu.release()
                xs.free_simulation_data(bmats)
            finally:
                os.environ.pop("XSB200_BANDS", None)
                os.environ.pop("XSB200_BAND_INDEX", None)

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        probe, kind, cores, _, _, variant = cpu_reference_rate(inp, sd, min(args.lookups, 300_000), args.kernel)
        sample = int(min(args.lookups, max(300_000, probe * 15.0)))
        rate, kind, cores, v, dt, variant = cpu_reference_rate(inp, sd, sample, args.kernel)
        cpu = {"value": rate, "unit": "lookups/s", "cores": cores, "kind": kind, "variant": variant,
               "sample": f"first {sample} of {args.lookups} lookups of the same workload, {dt:.1f} s",
               "checksum_of_sample": v % 999983}
    xs.free_simulation_data(sd)
    arith_ops = getattr(info, "fp64_ops_per_pair", 0) or 24
    gpu.release()

    # ---------------- reported baselines that need the GPU to themselves (rank 0, N=1 only) ----------------
    gpu_baseline = traffic = None
    if rank == 0 and world == 1:
        if not args.no_traffic_probe:
            traffic = measure_traffic(args, local_rank)
        if traffic is None:
            traffic = committed_traffic()
        if not args.no_gpu_baseline:
            gpu_baseline = gpu_reference_baseline(args.size, args.lookups, local_rank)

    if rank == 0:
        value = total_lookups * args.steps / elapsed_s
        kernel_s = statistics.mean(lookup_kernel_s)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        n_ueg = info.n_isotopes * info.n_gridpoints
        mean_nuc = pairs_mine / max(1, n_mine)
        alg = mean_nuc * (4 + 2 * 48) + math.ceil(math.log2(max(2, n_ueg))) * 8     # 5712.6 B at large
        alg_step = alg * args.lookups
        sm_mhz = (clock_info or {}).get("sm_mhz") or (clock_info or {}).get("sm_max_mhz") or 1965.0
        fp64_peak = info.sm_count * FP64_LANES_PER_CLK_PER_SM * sm_mhz * 1e6          # FP64 operations (lanes) per second
        k6_bytes = (traffic or {}).get("k6_lookup_phase_bytes")
        k0_bytes = (traffic or {}).get("k0_kernel_bytes")
        k4_bytes = (traffic or {}).get("k4_lookup_phase_bytes")
        hbm_view = lambda secs, dram: {   # noqa: E731
            "algorithmic_bytes_per_lookup": alg, "algorithmic_gbs": alg_step / secs / 1e9,
            "algorithmic_over_peak": alg_step / secs / 1e9 / hbm_peak,
            "traffic": dram, "traffic_over_algorithmic": (dram / alg_step) if dram else None,
            "hbm_achieved_gbs": (dram / secs / 1e9) if dram else None,
            "hbm_achieved_frac": (dram / secs / 1e9 / hbm_peak) if dram else None,
            "peak": hbm_peak, "peak_source": peak_src}
        if args.kernel == 6:
            floor_s = pairs_mine * arith_ops / fp64_peak
            roofline = {
                "bound": "fp64_issue", "achieved": pairs_mine * arith_ops / kernel_s / 1e12, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": floor_s / kernel_s,
                "unit_note": "FP64 operations = instruction lanes (an FMA counts once): the FP64 pipe issues 64 lanes/clk/SM whatever the opcode",
                "floor_ms": 1e3 * floor_s, "kernel_ms": 1e3 * kernel_s,
                "pairs": pairs_mine, "fp64_ops_per_pair": arith_ops,
                "arithmetic": "reference roundings (24 FP64 operations per (lookup, nuclide); macro_xs bit-identical)" if arith_ops == 24
                              else f"fused ({arith_ops} FP64 operations per (lookup, nuclide); macro_xs within 1e-12, integers guarded)",
                "peak_source": f"{info.sm_count} SMs x {FP64_LANES_PER_CLK_PER_SM} FP64 lanes/clk (scripts/exp/fp64_ops.cu) x {sm_mhz:.0f} MHz (sampled during the timed region)",
                "kernel": "lookup phase of -k 6 (CUDA events around it, mean over the timed steps): xs_dense_kernel<unionized> (materials with "
                          ">= 64 lookups per grid interval: 94 % of the lookups, 98.7 % of the pairs at 17 M) + xs_sorted_kernel<unionized> (the rest)",
                "frac_of_step": floor_s / (elapsed_s / args.steps),
                "traffic": k6_bytes, "traffic_source": (traffic or {}).get("source"),
                "hbm": hbm_view(kernel_s, k6_bytes),
                "why_not_hbm": "the energy-sorted lookups of a warp-group share their pair records (one fetch per 96 lookups) and two "
                               "index-row segments per group and chunk: measured DRAM traffic is a few % of the algorithmic gather bytes",
            }
        else:
            dram = k0_bytes if args.kernel < 4 else k4_bytes
            roofline = {"bound": "hbm", "achieved": alg_step / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_step / kernel_s / 1e9 / hbm_peak, "traffic": dram, "traffic_source": (traffic or {}).get("source"),
                        "kernel_ms": 1e3 * kernel_s, "peak_source": peak_src, "hbm": hbm_view(kernel_s, dram),
                        "kernel": "xs_tile_kernel<unionized>" if args.kernel < 4 else "xs_window_kernel<unionized> (all launches of one step)"}
        if "k0" in variants:
            secs = variants["k0"]["lookup_phase_ms"] * 1e-3
            # -k 0 is one launch that keeps a window of pair records L2-resident: DRAM moves a fraction of the
            # algorithmic bytes, so neither "algorithmic bytes / HBM peak" nor DRAM bandwidth is a fraction of a
            # roofline that binds; what does is the latency of the L2 gather (ncu: long-scoreboard stalls 47 %,
            # lts 40 %, l1tex 43 %: profiles/r02_tile_kernel_k0_ncu.txt).  Reported: both HBM views, no `frac`.
            variants["k0"]["roofline"] = dict(bound="l2 gather latency (window of pair records resident in L2)",
                                              algorithmic_gbs=alg_step / secs / 1e9, algorithmic_over_hbm_peak=alg_step / secs / 1e9 / hbm_peak,
                                              traffic=k0_bytes, traffic_over_algorithmic=(k0_bytes / alg_step) if k0_bytes else None,
                                              hbm_achieved_gbs=(k0_bytes / secs / 1e9) if k0_bytes else None,
                                              hbm_achieved_frac=(k0_bytes / secs / 1e9 / hbm_peak) if k0_bytes else None, hbm_peak=hbm_peak,
                                              kernel="xs_tile_kernel<unionized> (one launch: tiles grouped in shared memory, windowed sweep)")
        if "k6_fused" in variants:
            vf = variants["k6_fused"]
            f_floor = pairs_mine * vf["fp64_ops_per_pair"] / fp64_peak
            vf["roofline"] = dict(bound="fp64_issue", floor_ms=1e3 * f_floor, frac=f_floor / (vf["lookup_phase_ms"] * 1e-3),
                                  note="against its own floor (12 operations); not FP64-bound: shared-memory record delivery and issue slots "
                                       "bind first (profiles/r02_notes.md)")
        if "k4" in variants and k4_bytes:
            secs = variants["k4"]["lookup_phase_ms"] * 1e-3
            variants["k4"]["roofline"] = dict(bound="l1/l2 gather (see DESIGN.md 5.1b)", traffic=k4_bytes,
                                              traffic_over_algorithmic=k4_bytes / alg_step,
                                              hbm_achieved_frac=k4_bytes / secs / 1e9 / hbm_peak)
        total_inp = xs.read_CLI(["-s", args.size, "-m", "event", "-G", "unionized", "-l", str(total_lookups)])
        expected = xs.expected_checksum(total_inp)      # table covers 1, 2, 4, 8 x 17 M
        mean_phase = [1e3 * statistics.mean(p[i] for p in phase_s) for i in range(4)]
        line = {
            "metric": "XS lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "lookups_per_gpu": args.lookups,
                       "grid_bytes_per_gpu": int(info.resident_bytes), "parallelism": f"lookup-id sharding x{world}, grid replicated",
                       "l2": "no flush: 5.9 GB working set per GPU >> 126 MB L2", "timing": "cuda events, max over ranks",
                       "init": "host generator + upload" if world == 1 else "device-side generator per rank"},
            "checksum": verification % 999983, "checksum_expected": expected,
            "checksum_ok": (verification % 999983 == expected) if expected is not None else None,
            "lookups_counted": counted,
            "phase_ms": {"sample": mean_phase[0], "sort": mean_phase[1], "lookup": mean_phase[2], "reduce": mean_phase[3]},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_baseline": gpu_baseline,
            "e2e": {"value": total_lookups * e2e_steps / e2e_s, "unit": "lookups/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "device_ms_per_step": 1e3 * statistics.mean(e2e_device_s),
                    "h2d_gbs": h2d / statistics.mean(e2e_device_s) / 1e9 if e2e_device_s else None,
                    "checksum_matches_device_sampled": bool(e2e_ok),
                    "h2d_bytes_per_lookup": h2d / max(1, n_mine), "copy_only": copy_floor,
                    "host_side": ("the caller's int materials are narrowed to bytes by host threads inside the call, while the DMA engine "
                                  "moves the energies (xs_hostpack.h; XSB200_HOST_PACK=0 sends the ints: 12 B/lookup)") if h2d < 12 * n_mine
                                 else "samples copied as the caller holds them (8 B energy + 4 B material per lookup)",
                    "api": "xs_gpu_lookup_samples (pinned host energy/material samples -> checksum)"},
            "gpu_launches": launches, "clocks": clock_info, "variants": variants,
            "strong": strong, "bands": bands,
            "init": {"generate_s": round(t_gen, 2), "upload_s": round(t_up, 2)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
