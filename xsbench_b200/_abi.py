"""ctypes view of the two in-tree shared libraries.

* ``libxsb200.so``       -- CUDA kernels + the C ABI declared in ``include/xs_gpu.h``
* ``libxsb200_host.so``  -- host-side C (CLI, generator, report): ``host/xs_host.h``

The struct layouts mirror the reference's data model (cuda/XSbench_header.cuh:42-85):
``sizeof(Inputs) == 64``, ``sizeof(NuclideGridPoint) == 48``, ``sizeof(SimulationData) == 128``.
Nothing here computes anything: it only declares signatures.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
GPU_LIB_PATH = os.environ.get("XSB200_GPU_LIB") or os.path.join(PKG_DIR, "libxsb200.so")   # override: experiments only
HOST_LIB_PATH = os.path.join(PKG_DIR, "libxsb200_host.so")

# constants (include/xs_gpu.h)
UNIONIZED, NUCLIDE, HASH = 0, 1, 2
HISTORY_BASED, EVENT_BASED = 1, 2
HASH_MODULUS = 999983
XS_OK, XS_ERR_ARG, XS_ERR_CUDA, XS_ERR_UNSUPP, XS_ERR_NCCL = 0, -1, -2, -3, -4
N_PHASES = 4

#: every symbol include/xs_gpu.h declares
GPU_SYMBOLS = (
    "xs_gpu_init", "xs_gpu_run", "xs_gpu_run_range", "xs_gpu_lookup_samples", "xs_gpu_dump", "xs_gpu_sort_keys", "xs_gpu_read_array",
    "xs_gpu_selftest_division", "xs_gpu_narrow_materials", "xs_gpu_set_stream", "xs_gpu_finalize", "xs_gpu_get_info", "xs_gpu_last_error", "xs_gpu_version",
)
#: every symbol host/xs_host.h declares
HOST_SYMBOLS = (
    "LCG_random_double", "fast_forward_LCG", "pick_mat", "xs_material_thresholds", "xs_parse_cli",
    "read_CLI", "print_CLI_error", "xs_strip_driver_opts", "grid_init_do_not_profile", "xs_materials_only",
    "xs_free_simulation_data", "load_num_nucs", "load_mats", "load_concs", "NGP_compare",
    "double_compare", "estimate_mem_usage", "get_time", "logo", "center_print", "border_print",
    "fancy_int", "print_inputs", "print_results", "xs_expected_checksum", "binary_write", "binary_read",
)


class NuclideGridPoint(C.Structure):
    _fields_ = [("energy", C.c_double), ("total_xs", C.c_double), ("elastic_xs", C.c_double),
                ("absorbtion_xs", C.c_double), ("fission_xs", C.c_double), ("nu_fission_xs", C.c_double)]


class Inputs(C.Structure):
    _fields_ = [("nthreads", C.c_int), ("n_isotopes", C.c_long), ("n_gridpoints", C.c_long),
                ("lookups", C.c_int), ("HM", C.c_char_p), ("grid_type", C.c_int), ("hash_bins", C.c_int),
                ("particles", C.c_int), ("simulation_method", C.c_int), ("binary_mode", C.c_int),
                ("kernel_id", C.c_int)]


class SimulationData(C.Structure):
    _fields_ = [("num_nucs", C.POINTER(C.c_int)), ("concs", C.POINTER(C.c_double)),
                ("mats", C.POINTER(C.c_int)), ("unionized_energy_array", C.POINTER(C.c_double)),
                ("index_grid", C.POINTER(C.c_int)), ("nuclide_grid", C.POINTER(NuclideGridPoint)),
                ("length_num_nucs", C.c_int), ("length_concs", C.c_int), ("length_mats", C.c_int),
                ("length_unionized_energy_array", C.c_int), ("length_index_grid", C.c_long),
                ("length_nuclide_grid", C.c_int), ("max_num_nucs", C.c_int),
                ("verification", C.POINTER(C.c_ulong)), ("length_verification", C.c_int),
                ("p_energy_samples", C.POINTER(C.c_double)), ("length_p_energy_samples", C.c_int),
                ("mat_samples", C.POINTER(C.c_int)), ("length_mat_samples", C.c_int)]


class GpuResult(C.Structure):
    _fields_ = [("verification", C.c_ulonglong), ("n_lookups", C.c_ulonglong),
                ("device_seconds", C.c_double), ("phase_seconds", C.c_double * N_PHASES),
                ("host_seconds", C.c_double), ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong),
                ("gpu_launches", C.c_int), ("n_gpus", C.c_int)]


class GpuInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("l2_bytes", C.c_long),
                ("resident_bytes", C.c_long), ("n_isotopes", C.c_long), ("n_gridpoints", C.c_long),
                ("grid_type", C.c_int), ("hash_bins", C.c_int), ("max_num_nucs", C.c_int),
                ("n_ueg", C.c_long), ("fp64_ops_per_pair", C.c_int)]


class DriverOpts(C.Structure):
    _fields_ = [("gpus", C.c_int), ("reps", C.c_int), ("json", C.c_int), ("dump_xs", C.c_long), ("device_init", C.c_int)]


assert C.sizeof(Inputs) == 64 and C.sizeof(NuclideGridPoint) == 48 and C.sizeof(SimulationData) == 128

_gpu = None
_host = None


class ExtensionMissing(RuntimeError):
    """The in-tree CUDA extension is not built.  There is deliberately no fallback."""


def host_lib() -> C.CDLL:
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise ExtensionMissing(f"{HOST_LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(HOST_LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.LCG_random_double.restype = C.c_double
    lib.LCG_random_double.argtypes = [C.POINTER(C.c_uint64)]
    lib.fast_forward_LCG.restype = C.c_uint64
    lib.fast_forward_LCG.argtypes = [C.c_uint64, C.c_uint64]
    lib.pick_mat.restype = C.c_int
    lib.pick_mat.argtypes = [C.POINTER(C.c_uint64)]
    lib.xs_material_thresholds.restype = None
    lib.xs_material_thresholds.argtypes = [C.POINTER(C.c_double)]
    lib.xs_parse_cli.restype = C.c_int
    lib.xs_parse_cli.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(Inputs), C.c_char_p, C.c_size_t]
    lib.xs_strip_driver_opts.restype = C.c_int
    lib.xs_strip_driver_opts.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(DriverOpts)]
    lib.grid_init_do_not_profile.restype = SimulationData
    lib.grid_init_do_not_profile.argtypes = [Inputs, C.c_int]
    lib.xs_materials_only.restype = SimulationData
    lib.xs_materials_only.argtypes = [Inputs]
    lib.xs_free_simulation_data.restype = None
    lib.xs_free_simulation_data.argtypes = [C.POINTER(SimulationData)]
    lib.load_num_nucs.restype = C.POINTER(C.c_int)
    lib.load_num_nucs.argtypes = [C.c_long]
    lib.load_mats.restype = C.POINTER(C.c_int)
    lib.load_mats.argtypes = [C.POINTER(C.c_int), C.c_long, C.POINTER(C.c_int)]
    lib.load_concs.restype = C.POINTER(C.c_double)
    lib.load_concs.argtypes = [C.POINTER(C.c_int), C.c_int]
    lib.estimate_mem_usage.restype = C.c_size_t
    lib.estimate_mem_usage.argtypes = [Inputs]
    lib.print_inputs.restype = None
    lib.print_inputs.argtypes = [Inputs, C.c_int, C.c_int]
    lib.print_results.restype = C.c_int
    lib.print_results.argtypes = [Inputs, C.c_int, C.c_double, C.c_int, C.c_ulonglong]
    lib.xs_expected_checksum.restype = C.c_long
    lib.xs_expected_checksum.argtypes = [C.POINTER(Inputs)]
    lib.fancy_int.restype = None
    lib.fancy_int.argtypes = [C.c_long]
    lib.binary_write.restype = None
    lib.binary_write.argtypes = [Inputs, SimulationData]
    lib.binary_read.restype = SimulationData
    lib.binary_read.argtypes = [Inputs]
    _host = lib
    return lib


def gpu_lib() -> C.CDLL:
    """Load libxsb200.so.  Raises ExtensionMissing if it is not built -- never falls back."""
    global _gpu
    if _gpu is not None:
        return _gpu
    if not os.path.exists(GPU_LIB_PATH):
        raise ExtensionMissing(f"{GPU_LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(GPU_LIB_PATH, mode=C.RTLD_GLOBAL)
    ctx_p = C.c_void_p
    lib.xs_gpu_init.restype = C.c_int
    lib.xs_gpu_init.argtypes = [C.POINTER(Inputs), C.POINTER(SimulationData), C.c_int, C.POINTER(ctx_p)]
    lib.xs_gpu_run.restype = C.c_int
    lib.xs_gpu_run.argtypes = [ctx_p, C.POINTER(Inputs), C.POINTER(GpuResult)]
    lib.xs_gpu_run_range.restype = C.c_int
    lib.xs_gpu_run_range.argtypes = [ctx_p, C.POINTER(Inputs), C.c_long, C.c_long, C.POINTER(GpuResult)]
    lib.xs_gpu_lookup_samples.restype = C.c_int
    lib.xs_gpu_lookup_samples.argtypes = [ctx_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.POINTER(GpuResult)]
    lib.xs_gpu_dump.restype = C.c_int
    lib.xs_gpu_dump.argtypes = [ctx_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.xs_gpu_narrow_materials.restype = C.c_int
    lib.xs_gpu_narrow_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    lib.xs_gpu_selftest_division.restype = C.c_int
    lib.xs_gpu_selftest_division.argtypes = [ctx_p, C.c_ulonglong, C.c_long, C.c_int, C.POINTER(C.c_ulonglong)]
    lib.xs_gpu_sort_keys.restype = C.c_int
    lib.xs_gpu_sort_keys.argtypes = [ctx_p, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    lib.xs_gpu_read_array.restype = C.c_int
    lib.xs_gpu_read_array.argtypes = [ctx_p, C.c_int, C.c_long, C.c_long, C.c_void_p]
    lib.xs_gpu_set_stream.restype = C.c_int
    lib.xs_gpu_set_stream.argtypes = [ctx_p, C.c_void_p]
    lib.xs_gpu_finalize.restype = C.c_int
    lib.xs_gpu_finalize.argtypes = [ctx_p]
    lib.xs_gpu_get_info.restype = C.c_int
    lib.xs_gpu_get_info.argtypes = [ctx_p, C.POINTER(GpuInfo)]
    lib.xs_gpu_last_error.restype = C.c_char_p
    lib.xs_gpu_last_error.argtypes = []
    lib.xs_gpu_version.restype = C.c_char_p
    lib.xs_gpu_version.argtypes = []
    _gpu = lib
    return lib
