"""Python mirror of the host driver, for tests and bench.py.

The function names follow the reference's driver (cuda/Main.cu:3-109):

    in  = read_CLI(argv)                      cuda/io.cu:226
    SD  = grid_init_do_not_profile(in)        cuda/GridInit.cu:90
    gpu = move_simulation_data_to_device(in, SD)   cuda/GridInit.cu:4   -> xs_gpu_init
    res = gpu.run(in)                         run_event_based_simulation_*  -> xs_gpu_run
    gpu.release()                             release_device_memory    -> xs_gpu_finalize

Everything numeric happens inside the two C/CUDA libraries; this module only marshals.
There is no CPU implementation of the lookup here and no fallback: if libxsb200.so is
missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import (EVENT_BASED, HASH, HASH_MODULUS, HISTORY_BASED, NUCLIDE, UNIONIZED, GpuInfo,
                   GpuResult, Inputs, SimulationData)


class XSGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"xs_gpu error {code}: {message}")
        self.code = code


class CLIError(ValueError):
    """Usage error (the reference prints the usage text and exits with status 4)."""


def read_CLI(argv: Sequence[str]) -> Inputs:
    """Parse reference-style arguments (without the program name). Raises CLIError."""
    lib = _abi.host_lib()
    args = [b"xsbench"] + [a.encode() for a in argv]
    arr = (C.c_char_p * len(args))(*args)
    inp = Inputs()
    err = C.create_string_buffer(256)
    if lib.xs_parse_cli(len(args), arr, C.byref(inp), err, len(err)) != 0:
        raise CLIError(err.value.decode())
    # HM points into `arr`; keep it alive with the struct
    inp._keepalive = (arr, args)
    return inp


def make_inputs(size: str = "small", method: str = "event", grid: str = "unionized",
                lookups: Optional[int] = None, particles: Optional[int] = None,
                gridpoints: Optional[int] = None, hash_bins: Optional[int] = None,
                kernel_id: int = 0) -> Inputs:
    """Convenience wrapper building the argv the reference CLI would take."""
    argv = ["-s", size, "-m", method, "-G", grid, "-k", str(kernel_id)]
    if lookups is not None:
        argv += ["-l", str(lookups)]
    if particles is not None:
        argv += ["-p", str(particles)]
    if gridpoints is not None:
        argv += ["-g", str(gridpoints)]
    if hash_bins is not None:
        argv += ["-h", str(hash_bins)]
    return read_CLI(argv)


def grid_init_do_not_profile(inp: Inputs, mype: int = 1) -> SimulationData:
    """Generate the synthetic problem on the host (mype != 0 silences the progress lines)."""
    return _abi.host_lib().grid_init_do_not_profile(inp, mype)


def materials_only(inp: Inputs) -> SimulationData:
    """Material tables only; passing this to move_simulation_data_to_device builds the grids on the GPU."""
    return _abi.host_lib().xs_materials_only(inp)


def free_simulation_data(sd: SimulationData) -> None:
    _abi.host_lib().xs_free_simulation_data(C.byref(sd))


def simulation_arrays(inp: Inputs, sd: SimulationData) -> dict:
    """numpy views (no copy) of the six generated arrays."""
    def view(ptr, n, dtype):
        if n == 0 or not ptr:
            return np.empty(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_byte)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)
    out = {
        "num_nucs": view(sd.num_nucs, sd.length_num_nucs, np.int32),
        "concs": view(sd.concs, sd.length_concs, np.float64),
        "mats": view(sd.mats, sd.length_mats, np.int32),
        "nuclide_grid": view(sd.nuclide_grid, sd.length_nuclide_grid * 6, np.float64),
        "unionized_energy_array": view(sd.unionized_energy_array, sd.length_unionized_energy_array, np.float64),
        "index_grid": view(sd.index_grid, sd.length_index_grid, np.int32),
    }
    return out


@dataclasses.dataclass
class RunResult:
    verification: int            # un-modded
    n_lookups: int
    device_seconds: float
    phase_seconds: tuple
    host_seconds: float
    h2d_bytes: int
    d2h_bytes: int
    gpu_launches: int
    n_gpus: int

    @property
    def checksum(self) -> int:
        """The reference's final hash step (cuda/Main.cu:103)."""
        return self.verification % HASH_MODULUS

    @property
    def lookups_per_sec(self) -> float:
        return self.n_lookups / self.device_seconds if self.device_seconds > 0 else 0.0


def _result(r: GpuResult) -> RunResult:
    return RunResult(r.verification, r.n_lookups, r.device_seconds, tuple(r.phase_seconds),
                     r.host_seconds, r.h2d_bytes, r.d2h_bytes, r.gpu_launches, r.n_gpus)


class DeviceSimulation:
    """Owner of an ``xs_gpu_ctx`` (the device-resident problem)."""

    def __init__(self, inp: Inputs, sd: SimulationData, n_gpus: int = 1):
        self._lib = _abi.gpu_lib()
        self._ctx = C.c_void_p()
        self.inputs = inp
        rc = self._lib.xs_gpu_init(C.byref(inp), C.byref(sd), n_gpus, C.byref(self._ctx))
        if rc != _abi.XS_OK:
            self._ctx = C.c_void_p()
            raise XSGpuError(rc, self._lib.xs_gpu_last_error().decode())

    def _check(self, rc: int) -> None:
        if rc != _abi.XS_OK:
            raise XSGpuError(rc, self._lib.xs_gpu_last_error().decode())

    def run(self, inp: Optional[Inputs] = None) -> RunResult:
        res = GpuResult()
        self._check(self._lib.xs_gpu_run(self._ctx, C.byref(inp or self.inputs), C.byref(res)))
        return _result(res)

    def run_range(self, first_id: int, count: int, inp: Optional[Inputs] = None) -> RunResult:
        res = GpuResult()
        self._check(self._lib.xs_gpu_run_range(self._ctx, C.byref(inp or self.inputs), first_id, count, C.byref(res)))
        return _result(res)

    def lookup_samples(self, energy: np.ndarray, mat: np.ndarray, want_macro_xs: bool = False,
                       n: Optional[int] = None, energy_ptr: Optional[int] = None, mat_ptr: Optional[int] = None):
        """Macroscopic lookups on host-provided samples. Returns (RunResult, macro_xs | None).

        ``energy_ptr``/``mat_ptr`` let a caller pass raw (e.g. pinned) host addresses."""
        if energy_ptr is None:
            energy = np.ascontiguousarray(energy, dtype=np.float64)
            mat = np.ascontiguousarray(mat, dtype=np.int32)
            n = energy.shape[0]
            assert mat.shape[0] == n
            energy_ptr, mat_ptr = energy.ctypes.data, mat.ctypes.data
        macro = np.empty((n, 5), dtype=np.float64) if want_macro_xs else None
        res = GpuResult()
        self._check(self._lib.xs_gpu_lookup_samples(self._ctx, energy_ptr, mat_ptr, n,
                                                    macro.ctypes.data if want_macro_xs else None, C.byref(res)))
        return _result(res), macro

    def dump(self, first_id: int, n: int):
        """(energy[n], mat[n], macro_xs[n,5], argmax[n]) the device computes for event lookups."""
        e = np.empty(n, np.float64); m = np.empty(n, np.int32)
        x = np.empty((n, 5), np.float64); a = np.empty(n, np.int32)
        self._check(self._lib.xs_gpu_dump(self._ctx, first_id, n, e.ctypes.data, m.ctypes.data,
                                          x.ctypes.data, a.ctypes.data))
        return e, m, x, a

    def sort_keys(self, keys: np.ndarray, lo_bit: int = 0, hi_bit: int = 32) -> np.ndarray:
        """Permutation that sorts bits [lo_bit, hi_bit) of the u32 keys (stable), computed on the device."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        perm = np.empty(keys.shape[0], dtype=np.uint32)
        self._check(self._lib.xs_gpu_sort_keys(self._ctx, keys.ctypes.data, keys.shape[0], lo_bit, hi_bit, perm.ctypes.data))
        return perm

    def selftest_division(self, seed: int, n_pairs: int, mode: int = 0) -> int:
        """Bit mismatches between the kernels' Newton-Markstein quotient and IEEE division over n_pairs generated pairs."""
        bad = C.c_ulonglong(0)
        self._check(self._lib.xs_gpu_selftest_division(self._ctx, seed, n_pairs, mode, C.byref(bad)))
        return int(bad.value)

    def read_array(self, which: str) -> np.ndarray:
        """Download one device-resident problem array: 'nuclide_grid' (f64, 6 per point),
        'unionized_energy_array' (f64) or 'index_grid' (i32)."""
        info = self.info()
        pts = info.n_isotopes * info.n_gridpoints
        code, dtype, n = {"nuclide_grid": (0, np.float64, pts * 6), "unionized_energy_array": (1, np.float64, pts),
                          "index_grid": (2, np.int32, (pts if info.grid_type == UNIONIZED else info.hash_bins) * info.n_isotopes)}[which]
        out = np.empty(n, dtype=dtype)
        self._check(self._lib.xs_gpu_read_array(self._ctx, code, 0, out.nbytes, out.ctypes.data))
        return out

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._lib.xs_gpu_set_stream(self._ctx, cuda_stream))

    def info(self) -> GpuInfo:
        info = GpuInfo()
        self._check(self._lib.xs_gpu_get_info(self._ctx, C.byref(info)))
        return info

    def release(self) -> None:
        if self._ctx:
            self._lib.xs_gpu_finalize(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def move_simulation_data_to_device(inp: Inputs, sd: SimulationData, n_gpus: int = 1) -> DeviceSimulation:
    return DeviceSimulation(inp, sd, n_gpus)


def expected_checksum(inp: Inputs) -> Optional[int]:
    v = _abi.host_lib().xs_expected_checksum(C.byref(inp))
    return None if v < 0 else int(v)


__all__ = [
    "CLIError", "DeviceSimulation", "EVENT_BASED", "HASH", "HISTORY_BASED", "NUCLIDE", "RunResult",
    "UNIONIZED", "XSGpuError", "expected_checksum", "free_simulation_data", "grid_init_do_not_profile",
    "make_inputs", "materials_only", "move_simulation_data_to_device", "read_CLI", "simulation_arrays",
]
