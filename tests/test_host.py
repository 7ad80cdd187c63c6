"""Host-side logic of the product (no GPU needed): CLI, data model, generator, report, and
that both shared libraries load and export what their headers declare."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import xsbench_b200 as xs
from xsbench_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- ABI -------------------------------------------------------------------------------------
def declared_functions(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text))


def test_gpu_library_exports_every_declared_symbol():
    lib = _abi.gpu_lib()
    declared = {f for f in declared_functions(os.path.join(ROOT, "include", "xs_gpu.h")) if f.startswith("xs_gpu_")}
    assert declared == set(_abi.GPU_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.xs_gpu_version()


def test_host_library_exports_every_declared_symbol():
    lib = _abi.host_lib()
    declared = declared_functions(os.path.join(ROOT, "xsbench_b200", "host", "xs_host.h"))
    assert declared == set(_abi.HOST_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layout_matches_reference_header():
    # cuda/XSbench_header.cuh:42-85; offsets probed in SURVEY.md 8(b)
    assert C.sizeof(_abi.Inputs) == 64 and C.sizeof(_abi.NuclideGridPoint) == 48
    assert C.sizeof(_abi.SimulationData) == 128
    off = {f[0]: getattr(_abi.Inputs, f[0]).offset for f in _abi.Inputs._fields_}
    assert off == {"nthreads": 0, "n_isotopes": 8, "n_gridpoints": 16, "lookups": 24, "HM": 32, "grid_type": 40,
                   "hash_bins": 44, "particles": 48, "simulation_method": 52, "binary_mode": 56, "kernel_id": 60}
    # the CPU struct (openmp-threading) shares the prefix up to max_num_nucs
    for name in ("num_nucs", "concs", "mats", "unionized_energy_array", "index_grid", "nuclide_grid",
                 "length_num_nucs", "length_index_grid", "length_nuclide_grid", "max_num_nucs"):
        assert getattr(_abi.SimulationData, name).offset == getattr(ol.RefSimulationData, name).offset


def test_sm100a_cubin_is_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", _abi.GPU_LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


# ---- no GPU => loud failure, never a fallback -------------------------------------------------
def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_argument_errors_do_not_need_a_device():
    lib = _abi.gpu_lib()
    ctx = C.c_void_p()
    assert lib.xs_gpu_init(None, None, 1, C.byref(ctx)) == _abi.XS_ERR_ARG
    assert b"NULL" in lib.xs_gpu_last_error()
    inp = xs.make_inputs(size="small", lookups=10, gridpoints=50)
    sd = xs.grid_init_do_not_profile(inp)
    assert lib.xs_gpu_init(C.byref(inp), C.byref(sd), 0, C.byref(ctx)) == _abi.XS_ERR_ARG
    assert lib.xs_gpu_init(C.byref(inp), C.byref(sd), 9, C.byref(ctx)) == _abi.XS_ERR_ARG
    bad = xs.make_inputs(size="small", lookups=10, gridpoints=51)          # does not match sd
    assert lib.xs_gpu_init(C.byref(bad), C.byref(sd), 1, C.byref(ctx)) == _abi.XS_ERR_ARG
    res = _abi.GpuResult()
    assert lib.xs_gpu_run(None, C.byref(inp), C.byref(res)) == _abi.XS_ERR_ARG
    assert lib.xs_gpu_finalize(None) == _abi.XS_OK
    xs.free_simulation_data(sd)


@pytest.mark.parametrize("n,threads", [(0, 1), (1, 4), (63, 3), (65_535, 8), (65_536, 8), (1_000_003, 1), (1_000_003, 5), (1_000_003, 16)])
def test_host_side_material_narrowing(n, threads):
    """xs_gpu_lookup_samples sends the caller's int materials over PCIe as bytes, narrowed by a few host threads
    (csrc/xs_hostpack.h): every value in [0, 255] survives, anything else becomes 255 (= rejected on the device),
    whatever the thread count and however ragged the size."""
    import numpy as np
    lib = _abi.gpu_lib()
    rng = np.random.default_rng(n + threads)
    m = rng.integers(0, 12, n).astype(np.int32)
    if n > 10:
        m[rng.integers(0, n, 8)] = np.array([-1, 256, 255, 2**31 - 1, -2**31, 12, 4096 + 3, 65536], dtype=np.int32)
    out = np.full(n + 64, 0xAA, dtype=np.uint8)
    assert lib.xs_gpu_narrow_materials(m.ctypes.data, out.ctypes.data, n, threads) == _abi.XS_OK
    want = np.where((m >= 0) & (m <= 255), m, 255).astype(np.uint8)
    assert np.array_equal(out[:n], want) and np.all(out[n:] == 0xAA)
    assert lib.xs_gpu_narrow_materials(m.ctypes.data, out.ctypes.data, -1, threads) == _abi.XS_ERR_ARG
    assert lib.xs_gpu_narrow_materials(m.ctypes.data, out.ctypes.data, n, 0) == _abi.XS_ERR_ARG


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a machine without a GPU")
def test_fails_loudly_without_a_gpu():
    inp = xs.make_inputs(size="small", lookups=10, gridpoints=50)
    sd = xs.grid_init_do_not_profile(inp)
    with pytest.raises(xs.XSGpuError) as ei:
        xs.move_simulation_data_to_device(inp, sd)
    assert ei.value.code == _abi.XS_ERR_CUDA and "no CPU fallback" in str(ei.value)
    xs.free_simulation_data(sd)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing under xsbench_b200/ or include/ may mention it."""
    for base in ("xsbench_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle/" not in text and "xs_oracle" not in text and "libxsoracle" not in text \
                        and "libxsref" not in text and "import oracle" not in text, os.path.join(dirpath, f)


# ---- CLI (cuda/io.cu:226-441) -----------------------------------------------------------------
def test_cli_defaults():
    i = xs.read_CLI([])
    assert (i.simulation_method, i.n_isotopes, i.n_gridpoints, i.lookups, i.particles) == (xs.HISTORY_BASED, 355, 11303, 34, 500000)
    assert (i.grid_type, i.hash_bins, i.kernel_id, i.binary_mode, i.HM) == (xs.UNIONIZED, 10000, 0, 0, b"large")


def test_cli_event_default_lookups_and_order_quirk():
    i = xs.read_CLI(["-m", "event"])
    assert (i.lookups, i.particles) == (17_000_000, 0)
    assert xs.read_CLI(["-m", "event", "-l", "1000"]).lookups == 1000
    assert xs.read_CLI(["-l", "1000", "-m", "event"]).lookups == 1000
    # -p BEFORE "-m event" suppresses the 34 x 500000 product (cuda/io.cu:302-311)
    q = xs.read_CLI(["-p", "100", "-m", "event"])
    assert (q.lookups, q.particles) == (34, 100)
    # ... but after it, the product has already happened
    q = xs.read_CLI(["-m", "event", "-p", "100"])
    assert (q.lookups, q.particles) == (17_000_000, 100)


def test_cli_sizes():
    assert xs.read_CLI(["-s", "small"]).n_isotopes == 68
    assert xs.read_CLI(["-s", "small"]).n_gridpoints == 11303
    assert xs.read_CLI(["-s", "SMALL"]).n_isotopes == 68                 # validated case-insensitively
    assert xs.read_CLI(["-s", "XL"]).n_gridpoints == 238847
    assert xs.read_CLI(["-s", "XXL"]).n_gridpoints == 501578              # (long)(238847 * 2.1)
    assert xs.read_CLI(["-s", "XL", "-g", "777"]).n_gridpoints == 777     # -g wins
    assert xs.read_CLI(["-g", "777", "-s", "XXL"]).n_gridpoints == 777


def test_cli_grid_kernel_hash_binary():
    i = xs.read_CLI(["-G", "hash", "-h", "123", "-k", "6", "-b", "write", "-t", "3"])
    assert (i.grid_type, i.hash_bins, i.kernel_id, i.binary_mode, i.nthreads) == (xs.HASH, 123, 6, 2, 3)
    assert xs.read_CLI(["-G", "nuclide"]).grid_type == xs.NUCLIDE


@pytest.mark.parametrize("argv", [["-x", "1"], ["-m"], ["-m", "both"], ["-G", "tree"], ["-s", "medium"], ["-l", "0"],
                                  ["-g", "0"], ["-h", "0"], ["-b", "append"], ["event"], ["-t", "0"], ["-l"]])
def test_cli_errors(argv):
    with pytest.raises(xs.CLIError):
        xs.read_CLI(argv)


def test_cli_error_exit_status_is_4_like_the_reference():
    exe = os.path.join(ROOT, "xsbench_b200", "xsbench")
    p = subprocess.run([exe, "-G", "tree"], capture_output=True, text=True)
    assert p.returncode == 4 and "Usage:" in p.stdout


def test_driver_long_options_are_stripped():
    lib = _abi.host_lib()
    args = [b"xsbench", b"-m", b"event", b"--gpus", b"4", b"--json", b"-l", b"5", b"--reps", b"3", b"--dump-xs", b"2"]
    arr = (C.c_char_p * len(args))(*args)
    argc = C.c_int(len(args))
    o = _abi.DriverOpts()
    assert lib.xs_strip_driver_opts(C.byref(argc), arr, C.byref(o)) == 0
    assert (o.gpus, o.reps, o.json, o.dump_xs) == (4, 3, 1, 2)
    assert [arr[i] for i in range(argc.value)] == [b"xsbench", b"-m", b"event", b"-l", b"5"]
    bad = (C.c_char_p * 2)(b"xsbench", b"--frobnicate")
    argc = C.c_int(2)
    assert lib.xs_strip_driver_opts(C.byref(argc), bad, C.byref(o)) != 0


# ---- RNG / materials -------------------------------------------------------------------------
def test_lcg_matches_oracle():
    h, o = _abi.host_lib(), ol.oracle()
    for n in (0, 1, 2, 5, 10**6, 2**33 + 1, 2**62 + 3):
        for seed in (1070, 42, 1144900):
            assert h.fast_forward_LCG(seed, n) == o.xo_lcg_skip(seed, n)
    a, b = C.c_uint64(1070), C.c_uint64(1070)
    for _ in range(3000):
        assert h.LCG_random_double(C.byref(a)) == o.xo_lcg_next(C.byref(b)) and a.value == b.value
    a, b = C.c_uint64(5), C.c_uint64(5)
    for _ in range(5000):
        assert h.pick_mat(C.byref(a)) == o.xo_pick_mat(C.byref(b)) and a.value == b.value


def test_material_thresholds():
    thr = (C.c_double * 12)(); oth = (C.c_double * 12)()
    _abi.host_lib().xs_material_thresholds(thr); ol.oracle().xo_mat_thresholds(oth)
    assert list(thr) == list(oth)
    # values probed from the reference: fuel threshold 0, last 0.861...
    assert thr[0] == 0.0 and thr[1] == 0.052 and abs(thr[11] - 0.861) < 1e-15
    assert list(thr) == sorted(thr)


# ---- generator -------------------------------------------------------------------------------
CASES = [("small", 1000, "unionized", 10000), ("small", 1000, "hash", 500), ("small", 1000, "nuclide", 10000),
         ("large", 300, "unionized", 10000), ("large", 300, "hash", 37), ("small", 2, "unionized", 10000),
         ("small", 3, "hash", 1), ("small", 5000, "unionized", 10000)]


@pytest.mark.parametrize("size,n_gp,grid,hb", CASES)
def test_generator_byte_identical_to_oracle(size, n_gp, grid, hb):
    inp = xs.make_inputs(size=size, method="event", grid=grid, lookups=100, gridpoints=n_gp, hash_bins=hb)
    sd = xs.grid_init_do_not_profile(inp)
    arr = xs.simulation_arrays(inp, sd)
    p = ol.OracleProblem(inp.n_isotopes, n_gp, inp.grid_type, hb)
    assert np.array_equal(arr["nuclide_grid"], p.nuclide_grid)
    assert np.array_equal(arr["unionized_energy_array"], p.ueg)
    assert np.array_equal(arr["index_grid"], p.index_grid)
    assert np.array_equal(arr["num_nucs"], p.num_nucs) and sd.max_num_nucs == p.max_num_nucs
    w = sd.max_num_nucs
    for m in range(12):
        n = p.num_nucs[m]
        assert np.array_equal(arr["mats"][m * w:m * w + n], p.mats[m * w:m * w + n])
        assert np.array_equal(arr["concs"][m * w:m * w + n], p.concs[m * w:m * w + n])
    assert sd.length_nuclide_grid == inp.n_isotopes * n_gp
    assert sd.length_concs == sd.length_mats == 12 * w and sd.length_num_nucs == 12
    xs.free_simulation_data(sd)


@pytest.mark.skipif(not ol.have_reference(), reason="oracle/_ref/libxsref.so not built")
@pytest.mark.parametrize("size,n_gp,grid,hb", CASES[:5] + CASES[7:])
def test_generator_byte_identical_to_reference(size, n_gp, grid, hb):
    inp = xs.make_inputs(size=size, method="event", grid=grid, lookups=100, gridpoints=n_gp, hash_bins=hb)
    sd = xs.grid_init_do_not_profile(inp)
    arr = xs.simulation_arrays(inp, sd)
    rinp = ol.ref_inputs(inp.n_isotopes, n_gp, inp.grid_type, hb)
    rsd = ol.reference().grid_init_do_not_profile(rinp, 1)
    n_pts = inp.n_isotopes * n_gp
    assert np.array_equal(arr["nuclide_grid"], np.ctypeslib.as_array(rsd.nuclide_grid, shape=(n_pts * 6,)))
    if rsd.length_index_grid:
        assert np.array_equal(arr["index_grid"], np.ctypeslib.as_array(rsd.index_grid, shape=(rsd.length_index_grid,)))
    if inp.grid_type == 0:
        assert np.array_equal(arr["unionized_energy_array"], np.ctypeslib.as_array(rsd.unionized_energy_array, shape=(n_pts,)))
    xs.free_simulation_data(sd)


def test_generator_is_thread_count_invariant():
    a = xs.read_CLI(["-s", "small", "-g", "3000", "-t", "1"])
    b = xs.read_CLI(["-s", "small", "-g", "3000", "-t", "7"])
    sa, sb = xs.grid_init_do_not_profile(a), xs.grid_init_do_not_profile(b)
    for k, v in xs.simulation_arrays(a, sa).items():
        assert np.array_equal(v, xs.simulation_arrays(b, sb)[k]), k
    xs.free_simulation_data(sa); xs.free_simulation_data(sb)


def test_estimate_mem_usage_matches_reference_formula():
    i = xs.read_CLI(["-s", "large"])
    assert _abi.host_lib().estimate_mem_usage(i) == 5649 + 0 or _abi.host_lib().estimate_mem_usage(i) > 5600
    pts = 355 * 11303
    want = int(np.ceil((pts * 48 + pts * 8 + pts * 355 * 4) / 1048576.0))
    assert _abi.host_lib().estimate_mem_usage(i) == want


# ---- report ----------------------------------------------------------------------------------
def test_expected_checksum_table():
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-s", "small"])) == 945990
    assert xs.expected_checksum(xs.read_CLI(["-m", "event"])) == 952131
    assert xs.expected_checksum(xs.read_CLI(["-s", "small"])) == 941535
    assert xs.expected_checksum(xs.read_CLI([])) == 954318
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-s", "small", "-g", "1000", "-l", "100000"])) == 302880
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-s", "XL", "-l", "1000000"])) == 3377
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-s", "XXL"])) == 33803
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-l", "12345"])) is None
    # 10^9 lookups: the reference CUDA build's printed value corrected for its 32-bit accumulator
    # (thrust::reduce(..., 0), cuda/Simulation.cu:34): sums above 2^31 lose 2^64 - 2^32
    wrap = (2**64 - 2**32) % 999983
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-l", "1000000000"])) == (260078 - wrap) % 999983 == 296043
    assert xs.expected_checksum(xs.read_CLI(["-m", "event", "-s", "XL", "-l", "1000000000"])) == (509518 - wrap) % 999983 == 545483


def test_print_results_validity(capfd):
    lib = _abi.host_lib()
    i = xs.read_CLI(["-m", "event", "-s", "small"])
    assert lib.print_results(i, 0, 0.5, 1, 945990) == 0
    assert lib.print_results(i, 0, 0.5, 1, 945991) == 1
    out = capfd.readouterr().out
    assert "Verification checksum: 945990 (Valid)" in out and "INVALID CHECKSUM" in out
    assert "Lookups/s:   34,000,000" in out


def test_binary_roundtrip(tmp_path):
    exe_cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        lib = _abi.host_lib()
        inp = xs.read_CLI(["-s", "small", "-g", "200", "-G", "hash", "-h", "50"])
        sd = xs.grid_init_do_not_profile(inp)
        lib.binary_write(inp, sd)
        back = lib.binary_read(inp)
        a, b = xs.simulation_arrays(inp, sd), xs.simulation_arrays(inp, back)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert back.max_num_nucs == sd.max_num_nucs
        xs.free_simulation_data(sd); xs.free_simulation_data(back)
    finally:
        os.chdir(exe_cwd)


def test_binary_read_accepts_reference_files(tmp_path):
    """`-b read` falls back to the reference's own format when the header magic is absent: a raw
    SimulationData dump (128 bytes for the cuda/ port, 112 for openmp-threading; stale pointers
    included) followed by the six arrays (cuda/io.cu:443-495)."""
    import ctypes as C
    exe_cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        lib = _abi.host_lib()
        for argv in (["-s", "small", "-g", "150", "-G", "unionized"], ["-s", "small", "-g", "150", "-G", "hash", "-h", "40"],
                     ["-s", "small", "-g", "150", "-G", "nuclide"]):
            inp = xs.read_CLI(argv)
            sd = xs.grid_init_do_not_profile(inp)
            a = xs.simulation_arrays(inp, sd)
            for struct_bytes in (128, 112):
                raw = bytes(sd)                              # our struct is laid out like the cuda/ port's (128 bytes)
                assert len(raw) == 128
                with open("XS_data.dat", "wb") as f:
                    f.write(raw[:80] + b"\xaa" * (struct_bytes - 80))      # shared prefix + that port's stale tail
                    for k in ("num_nucs", "concs", "mats", "nuclide_grid", "index_grid", "unionized_energy_array"):
                        f.write(a[k].tobytes())
                back = lib.binary_read(inp)
                b = xs.simulation_arrays(inp, back)
                for k in a:
                    assert np.array_equal(a[k], b[k]), (argv, struct_bytes, k)
                assert back.max_num_nucs == sd.max_num_nucs
                xs.free_simulation_data(back)
            xs.free_simulation_data(sd)
    finally:
        os.chdir(exe_cwd)


def test_binary_read_of_a_file_the_reference_wrote(tmp_path):
    """The unmodified reference (openmp-threading, oracle/_ref/XSBench_ref) writes XS_data.dat; our
    binary_read loads it and the arrays equal our own generator's."""
    import subprocess
    import oracle_lib as ol
    if not os.path.exists(ol.REF_BIN):
        pytest.skip("oracle/_ref/XSBench_ref not built")
    exe_cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        subprocess.run([ol.REF_BIN, "-s", "small", "-g", "120", "-m", "event", "-l", "100", "-b", "write", "-t", "2"],
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
        assert os.path.exists("XS_data.dat")
        inp = xs.read_CLI(["-s", "small", "-g", "120"])
        back = _abi.host_lib().binary_read(inp)
        sd = xs.grid_init_do_not_profile(inp)
        a, b = xs.simulation_arrays(inp, sd), xs.simulation_arrays(inp, back)
        assert back.max_num_nucs == sd.max_num_nucs
        w = sd.max_num_nucs
        # (the reference leaves the unused tail of every material's row uninitialised)
        valid = np.concatenate([np.arange(m * w, m * w + a["num_nucs"][m]) for m in range(12)])
        for k in a:
            if k in ("concs", "mats"):
                assert np.array_equal(a[k][valid], b[k][valid]), k
            else:
                assert np.array_equal(a[k], b[k]), k
        xs.free_simulation_data(sd); xs.free_simulation_data(back)
    finally:
        os.chdir(exe_cwd)
