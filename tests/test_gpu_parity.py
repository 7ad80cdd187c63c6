"""Parity of the CUDA path against the CPU oracle, through the C ABI (include/xs_gpu.h).

Bars (BASELINE.json north_star): verification sums BIT-EXACT (integers derived from argmax);
macro_xs vectors within 1e-12 relative of the oracle (summation order differs: warp
reductions); sampled energies/materials bit-exact.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import xsbench_b200 as xs
from xsbench_b200 import _abi

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12          # north_star tolerance for macro_xs
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
GRIDS = {"unionized": 0, "nuclide": 1, "hash": 2}
NTHREADS = os.cpu_count() or 1


def unhex(xs_):
    return np.array([float.fromhex(x) for x in xs_])


class Problem:
    def __init__(self, size="small", n_gp=1000, grid="unionized", hb=500, **kw):
        self.inp = xs.make_inputs(size=size, grid=grid, gridpoints=n_gp, hash_bins=hb, **kw)
        self.sd = xs.grid_init_do_not_profile(self.inp)
        self.gpu = xs.move_simulation_data_to_device(self.inp, self.sd)
        self.oracle = ol.OracleProblem(self.inp.n_isotopes, n_gp, GRIDS[grid], hb)

    def close(self):
        self.gpu.release()
        xs.free_simulation_data(self.sd)
        self.oracle.close()


# Every grid type twice: with the default split of the sorted pipeline (at these sizes no material
# has 64 lookups per grid interval, so xs_sorted_kernel does everything) and with XSB200_DENSE_MIN=1
# (every material through xs_dense_kernel, sparse ones included: its per-lookup fallback runs a lot).
# The knob is read once, by xs_gpu_init.
@pytest.fixture(scope="module", params=[(g, dm) for g in ("unionized", "hash", "nuclide") for dm in (None, "1")],
                ids=lambda p: p[0] + ("-dense" if p[1] else ""))
def small(request):
    grid, dense_min = request.param
    if dense_min:
        os.environ["XSB200_DENSE_MIN"] = dense_min
    try:
        p = Problem("small", 1000, grid, 500, method="event", lookups=100000)
    finally:
        os.environ.pop("XSB200_DENSE_MIN", None)
    yield p
    p.close()


def max_rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


# ---- event mode, every kernel variant -----------------------------------------------------------
@pytest.mark.parametrize("kernel_id", [0, 1, 2, 3, 4, 5, 6])
def test_event_checksum_all_variants(small, kernel_id):
    inp = xs.make_inputs(size="small", grid={0: "unionized", 1: "nuclide", 2: "hash"}[small.inp.grid_type],
                         gridpoints=1000, hash_bins=500, method="event", lookups=100000, kernel_id=kernel_id)
    res = small.gpu.run(inp)
    assert res.n_lookups == 100000
    assert res.verification == 302880                      # golden, from the unmodified reference
    assert res.verification == small.oracle.event(0, 100000, NTHREADS)
    assert res.gpu_launches >= 1 and res.device_seconds > 0


def test_event_per_lookup_parity(small):
    n = 20000
    e, m, macro, am = small.gpu.dump(0, n)
    v, oe, om, omacro, oam = small.oracle.event_dump(0, n)
    assert np.array_equal(e, oe), "sampled energies must be bit-exact"
    assert np.array_equal(m, om), "sampled materials must be bit-exact"
    assert np.array_equal(am, oam), "argmax indices must be bit-exact"
    assert max_rel(macro, omacro) <= REL_TOL
    assert set(np.unique(m)) == set(range(12))


def test_event_golden_rows(small):
    """Committed vectors from the unmodified reference (incl. ids far into the stream)."""
    for g in GOLDEN["lookups"]:
        if (g["n_isotopes"], g["n_gridpoints"], g["grid_type"]) != (68, 1000, small.inp.grid_type):
            continue
        for row in g["rows"]:
            e, m, macro, am = small.gpu.dump(row["id"], 1)
            assert e[0] == float.fromhex(row["energy"]) and m[0] == row["mat"] and am[0] == row["argmax"]
            assert max_rel(macro[0], unhex(row["macro_xs"])) <= REL_TOL


def test_run_range_partition_is_exact(small):
    whole = small.gpu.run_range(0, 50001).verification
    parts = [small.gpu.run_range(a, b - a) for a, b in ((0, 1), (1, 32), (32, 33), (33, 20000), (20000, 50001))]
    assert sum(p.verification for p in parts) == whole == small.oracle.event(0, 50001, NTHREADS)
    assert sum(p.n_lookups for p in parts) == 50001
    far = small.gpu.run_range(900_000_000, 1000)           # ids near 2^30: 64-bit ids on the device
    assert far.verification == small.oracle.event(900_000_000, 1000, NTHREADS)
    empty = small.gpu.run_range(5, 0)
    assert empty.verification == 0 and empty.n_lookups == 0
    for k in (1, 4, 6):
        grid = {0: "unionized", 1: "nuclide", 2: "hash"}[small.inp.grid_type]
        inp_k = xs.make_inputs(size="small", grid=grid, gridpoints=1000, hash_bins=500, method="event", lookups=10, kernel_id=k)
        assert small.gpu.run_range(7, 0, inp_k).n_lookups == 0
        one = small.gpu.run_range(7, 1, inp_k)
        assert one.n_lookups == 1 and one.verification == small.oracle.event(7, 1, 1)


# ---- host-sample entry point (edge cases) ----------------------------------------------------------
def test_lookup_samples_matches_oracle(small):
    rng = np.random.default_rng(11)
    grid_e = small.oracle.nuclide_grid[0::6]
    e = np.concatenate([rng.random(5000),
                        [0.0, 5e-324, 1e-300, 1e-12, 0.5, 1.0 - 2.0**-53],        # range extremes
                        grid_e[:40], np.nextafter(grid_e[:40], 0), np.nextafter(grid_e[:40], 1),   # on / around grid points
                        [grid_e.min(), grid_e.max(), np.nextafter(grid_e.max(), 1)]])
    m = rng.integers(0, 12, len(e)).astype(np.int32)
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m)
    assert res.verification == v and res.n_lookups == len(e)
    assert max_rel(macro, omacro) <= REL_TOL
    # (8 B energy + the material narrowed to 1 byte on the host; XSB200_HOST_PACK=0: the caller's 4-byte int)
    # (on a host with fewer than 8 hardware threads per visible GPU the narrowing is off by default: 12 B)
    assert res.h2d_bytes in (len(e) * 9, len(e) * 12) and res.d2h_bytes >= len(e) * 40


def test_sweep_path_is_bit_identical_to_reference_order(small):
    """xs_gpu_lookup_samples runs the grouped, windowed sweep (same kernels as -k 4/5/6).  It adds
    the per-nuclide terms in the reference's order with the reference's roundings (division by
    Newton-Markstein correction of a stored reciprocal = correctly rounded quotient), so the
    vectors must equal the oracle's bit for bit, not just within 1e-12."""
    rng = np.random.default_rng(5)
    n = 200_000
    e = rng.random(n); m = rng.integers(0, 12, n).astype(np.int32)
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m)
    assert res.verification == v and res.n_lookups == n
    assert np.array_equal(macro, omacro)


def test_division_is_exact(small):
    """Stress the reciprocal-based division: energies ON grid points (f = 0 or 1), one ulp
    around them, and at the ends of every nuclide's grid."""
    g = small.oracle.nuclide_grid[0::6]
    e = np.concatenate([g[::7], np.nextafter(g[::7], 0), np.nextafter(g[::7], 1)])
    e = e[(e > 0) & (e < 1)]
    rng = np.random.default_rng(9)
    m = rng.integers(0, 12, len(e)).astype(np.int32)
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m)
    assert res.verification == v
    assert np.array_equal(macro, omacro)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_newton_markstein_quotient_equals_ieee_division(small, mode):
    """The lane-per-lookup kernels form f = (hi.E - E) / d from a stored, correctly rounded 1/d by one Newton-Markstein
    step and rely on the result being the correctly rounded quotient (macro_xs bit-identical to the reference, no
    near-tie guard on that path).  Markstein's theorem needs a faithful first approximation, which RN(n * RN(1/d)) is
    not guaranteed to be in general -- so the claim is tested, not assumed: 2^30 pairs as the lookups form them
    (lo < E <= hi in [0, 1)), 2^30 with random mantissas and exponents down to 2^-63, 2^28 adversarial divisors."""
    n = 1 << (28 if mode == 2 else 30)
    for seed in (1, 0x9e3779b97f4a7c15):
        assert small.gpu.selftest_division(seed, n, mode) == 0


def test_lookups_on_grid_points_only(small):
    """Every lookup energy IS a grid point of one of the first nuclides, in every material: each
    lookup sits exactly ON an interval bound of some nuclide, where "<" and "<=" part ways (and
    where the reference's hash-grid guards `E <= e_low` / `E >= e_high`, cuda/Simulation.cu:150-156,
    come into play).  The dense kernel sends such a lookup through the reference's own procedure;
    results must equal the oracle's bit for bit on every grid type."""
    n_gp = 1000
    grid_e = small.oracle.nuclide_grid[0::6]
    e1 = np.concatenate([grid_e[j * n_gp:(j + 1) * n_gp] for j in range(4)])
    e1 = e1[(e1 >= 0) & (e1 < 1)]
    e = np.tile(e1, 12)
    m = np.repeat(np.arange(12, dtype=np.int32), len(e1))
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m)
    assert res.verification == v and res.n_lookups == len(e)
    assert np.array_equal(macro, omacro)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000])
def test_lookup_samples_ragged_sizes(small, n):
    rng = np.random.default_rng(n)
    e, m = rng.random(n), rng.integers(0, 12, n).astype(np.int32)
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m) if n else (0, np.empty((0, 5)))
    assert res.verification == v and res.n_lookups == n
    if n:
        assert max_rel(macro, omacro) <= REL_TOL


@pytest.mark.parametrize("mat", [0, 2, 4, 11])
def test_lookup_samples_single_material(small, mat):
    rng = np.random.default_rng(mat)
    e = rng.random(3000); m = np.full(3000, mat, np.int32)
    res, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
    v, omacro = small.oracle.lookup_samples(e, m)
    assert res.verification == v and max_rel(macro, omacro) <= REL_TOL


def test_results_are_reproducible_run_to_run(small):
    a = small.gpu.dump(0, 4096)[2]
    b = small.gpu.dump(0, 4096)[2]
    assert np.array_equal(a, b)          # fixed reduction tree => bitwise deterministic


def test_sorted_pipeline_is_repeatable_on_one_context(small):
    """The lane-per-lookup kernels take their warp-groups from an atomic counter that the last
    warp of a launch re-arms.  Back-to-back runs on one context (-k 6 and the host-sample entry
    point, interleaved) must keep giving the same integers and bit-identical vectors: which warp
    gets which group changes from run to run, what a lookup computes does not."""
    grid = {0: "unionized", 1: "nuclide", 2: "hash"}[small.inp.grid_type]
    inp = xs.make_inputs(size="small", grid=grid, gridpoints=1000, hash_bins=500, method="event", lookups=100000, kernel_id=6)
    rng = np.random.default_rng(77)
    e = rng.random(30_011); m = rng.integers(0, 12, len(e)).astype(np.int32)
    first_macro = None
    for _ in range(4):
        res = small.gpu.run(inp)
        assert res.verification == 302880 and res.n_lookups == 100000
        r2, macro = small.gpu.lookup_samples(e, m, want_macro_xs=True)
        assert r2.n_lookups == len(e)
        if first_macro is None:
            first_macro, first_v = macro.copy(), r2.verification
        assert r2.verification == first_v and np.array_equal(macro, first_macro)


# ---- history mode (CPU-only in the reference: openmp-threading/Simulation.c:116-238) -------------------
@pytest.mark.parametrize("grid,hb", [("unionized", 10000), ("hash", 500), ("nuclide", 10000)])
def test_history_mode(grid, hb):
    p = Problem("small", 1000, grid, hb, method="history", lookups=34, particles=3000)
    try:
        res = p.gpu.run()
        assert res.n_lookups == 34 * 3000
        assert res.verification == 309181                   # golden (reference), grid-type invariant
        assert res.verification == p.oracle.history(0, 3000, 34, NTHREADS)
        part = p.gpu.run_range(1000, 500)
        assert part.verification == p.oracle.history(1000, 500, 34, NTHREADS)
        short = xs.make_inputs(size="small", grid=grid, gridpoints=1000, hash_bins=hb, method="history", lookups=7, particles=900)
        assert p.gpu.run(short).verification == p.oracle.history(0, 900, 7, NTHREADS)
    finally:
        p.close()


# ---- the large fuel (321 nuclides) at reduced grid size -------------------------------------------
@pytest.mark.parametrize("grid,hb", [("unionized", 10000), ("hash", 2000)])
def test_large_fuel_small_grid(grid, hb):
    p = Problem("large", 1000, grid, hb, method="event", lookups=100000)
    try:
        assert p.gpu.run().verification == 303045           # golden (reference)
        for k in (4, 6):
            inp = xs.make_inputs(size="large", grid=grid, gridpoints=1000, hash_bins=hb, method="event", lookups=100000, kernel_id=k)
            assert p.gpu.run(inp).verification == 303045
        e, m, macro, am = p.gpu.dump(0, 5000)
        _, oe, om, omacro, oam = p.oracle.event_dump(0, 5000)
        assert np.array_equal(e, oe) and np.array_equal(m, om) and np.array_equal(am, oam)
        assert max_rel(macro, omacro) <= REL_TOL
        for g in GOLDEN["lookups"]:
            if (g["n_isotopes"], g["n_gridpoints"], g["grid_type"]) == (355, 1000, p.inp.grid_type):
                for row in g["rows"]:
                    _, _, x, a = p.gpu.dump(row["id"], 1)
                    assert a[0] == row["argmax"] and max_rel(x[0], unhex(row["macro_xs"])) <= REL_TOL
    finally:
        p.close()


# ---- both gather mappings give the same integers -----------------------------------------------------
def test_gather_variants_agree(monkeypatch):
    sums = []
    for gather in ("0", "1"):
        monkeypatch.setenv("XSB200_GATHER", gather)
        p = Problem("small", 1000, "unionized", 500, method="event", lookups=50000)
        try:
            sums.append(p.gpu.run().verification)
            e, m, macro, am = p.gpu.dump(0, 3000)
            assert max_rel(macro, p.oracle.event_dump(0, 3000)[3]) <= REL_TOL
        finally:
            p.close()
    assert sums[0] == sums[1] == ol.OracleProblem(68, 1000, 0, 500).event(0, 50000, NTHREADS)


# ---- every way through the -k 6 pipeline gives the same integers -----------------------------------------
@pytest.mark.parametrize("env", [
    {"XSB200_BIN_BITS": "14"},          # one-pass bin sort instead of the radix sort
    {"XSB200_FUSE_GATHER": "0"},        # separate gather pass instead of indirect reads in the kernel
    {"XSB200_SORTED_KERNEL": "0"},      # windowed sweep on the sorted batch
    {"XSB200_KEY_LO_BIT": "26"},        # barely sorted: groups span many grid intervals -> direct-load path
    {"XSB200_E2E_KERNEL": "4"},         # host-sample API through partition + windowed sweep
    {"XSB200_DENSE_MIN": "0"},          # no dense kernel at all
    {"XSB200_DENSE_MIN": "20"},         # some materials dense, some not: two launches
    {"XSB200_DENSE_MIN": "1", "XSB200_KEY_LO_BIT": "26"},   # dense kernel on a barely sorted batch: nearly every lookup resolves itself
    {"XSB200_DENSE_MIN": "1", "XSB200_FUSE_GATHER": "0"},   # dense kernel on gathered (not indirect) samples
    {"XSB200_FUSE_DIGITS": "0"},        # the sort counts its digits in a pass of its own instead of in the sampler / locate kernels
    {"XSB200_ONESWEEP": "0"},           # round 1's three-kernel radix passes
    {"XSB200_HOST_PACK": "0"},          # host-sample API: materials cross PCIe as the caller's ints, not narrowed to bytes
    {"XSB200_HOST_PACK": "1", "XSB200_PACK_THREADS": "3"},   # ... narrowed by 3 host threads whatever the host looks like
])
@pytest.mark.parametrize("grid,hb", [("unionized", 500), ("hash", 500), ("nuclide", 500)])
def test_sorted_pipeline_variants_agree(monkeypatch, env, grid, hb):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    p = Problem("small", 1000, grid, hb, method="event", lookups=100000)
    try:
        inp = xs.make_inputs(size="small", grid=grid, gridpoints=1000, hash_bins=hb, method="event", lookups=100000, kernel_id=6)
        assert p.gpu.run(inp).verification == 302880
        rng = np.random.default_rng(3)
        n = 70_001
        e = rng.random(n); m = rng.integers(0, 12, n).astype(np.int32)
        res, macro = p.gpu.lookup_samples(e, m, want_macro_xs=True)
        v, omacro = p.oracle.lookup_samples(e, m)
        assert res.verification == v and np.array_equal(macro, omacro)
    finally:
        p.close()


@pytest.mark.parametrize("grid,hb", [("unionized", 50), ("hash", 50), ("nuclide", 50)])
def test_dense_kernel_at_the_default_split(grid, hb):
    """100 grid points per nuclide and 10^6 lookups: every material has >= 64 lookups per grid
    interval, so the default split sends all of them to xs_dense_kernel (as large/fuel at 17 M).
    Energies on grid points, one ulp around them and beyond every nuclide's ends are mixed in;
    macro_xs must equal the oracle's bit for bit."""
    p = Problem("small", 100, grid, hb, method="event", lookups=1_000_000)
    try:
        inp = xs.make_inputs(size="small", grid=grid, gridpoints=100, hash_bins=hb, method="event", lookups=1_000_000, kernel_id=6)
        assert p.gpu.run(inp).verification == p.oracle.event(0, 1_000_000, NTHREADS)
        rng = np.random.default_rng(17)
        g = p.oracle.nuclide_grid[0::6]
        e = np.concatenate([rng.random(400_000), g, np.nextafter(g, 0), np.nextafter(g, 1), [0.0, 5e-324, 1.0 - 2.0**-53]])
        e = e[(e >= 0) & (e < 1)]
        m = rng.integers(0, 12, len(e)).astype(np.int32)
        res, macro = p.gpu.lookup_samples(e, m, want_macro_xs=True)
        v, omacro = p.oracle.lookup_samples(e, m)
        assert res.verification == v and res.n_lookups == len(e)
        assert np.array_equal(macro, omacro)
    finally:
        p.close()


# ---- energy-band sharding of the unionized grid (what XXL needs), exercised on one GPU ----------------------
@pytest.mark.parametrize("device_init", [False, True])
def test_energy_bands_sum_to_the_whole(monkeypatch, device_init):
    """Each context holds the index rows of one energy band, draws EVERY lookup id and keeps the ones
    whose row lies in its band (SURVEY 8e option 2): the per-band sums add up to the full result,
    whatever -k variant is asked for."""
    n, bands = 100000, 3
    total_v = total_n = 0
    for b in range(bands):
        monkeypatch.setenv("XSB200_BANDS", str(bands))
        monkeypatch.setenv("XSB200_BAND_INDEX", str(b))
        inp = xs.make_inputs(size="small", method="event", grid="unionized", lookups=n, gridpoints=1000, kernel_id=6)
        sd = xs.materials_only(inp) if device_init else xs.grid_init_do_not_profile(inp)
        with xs.move_simulation_data_to_device(inp, sd) as gpu:
            r6 = gpu.run(inp)
            r0 = gpu.run(xs.make_inputs(size="small", method="event", grid="unionized", lookups=n, gridpoints=1000, kernel_id=0))
            assert (r0.verification, r0.n_lookups) == (r6.verification, r6.n_lookups)
            assert 0 < r6.n_lookups < n
            # the band's index rows are the corresponding rows of the full grid
            if not device_init:
                info = gpu.info()
                n_ueg = info.n_isotopes * info.n_gridpoints
                r_lo, r_hi = n_ueg * b // bands, n_ueg * (b + 1) // bands
                rows = np.empty((r_hi - r_lo) * info.n_isotopes, np.int32)
                gpu._check(gpu._lib.xs_gpu_read_array(gpu._ctx, 2, r_lo * info.n_isotopes * 4, rows.nbytes, rows.ctypes.data))
                full = xs.simulation_arrays(inp, sd)["index_grid"]
                assert np.array_equal(rows, full[r_lo * info.n_isotopes:r_hi * info.n_isotopes])
            # history mode needs all bands in one context (they exchange the particles' feedback each generation)
            with pytest.raises(xs.XSGpuError):
                gpu.run(xs.make_inputs(size="small", method="history", grid="unionized", gridpoints=1000, particles=100, lookups=5))
            total_v += r6.verification
            total_n += r6.n_lookups
        xs.free_simulation_data(sd)
    assert total_n == n and total_v == 302880


def test_energy_bands_host_samples_and_dump(monkeypatch):
    """xs_gpu_lookup_samples and xs_gpu_dump on band-sharded contexts: every band takes all samples, performs
    the lookups of its rows and leaves zeros elsewhere; the bands' checksums, counts and macro_xs rows add up
    to the un-sharded result (bit for bit: a sum with zeros)."""
    bands = 3
    rng = np.random.default_rng(29)
    e = rng.random(40_000); m = rng.integers(0, 12, len(e)).astype(np.int32)
    orc = ol.OracleProblem(68, 1000, 0, 10000)
    want_v, want_macro = orc.lookup_samples(e, m)
    _, oe, om, omacro, oam = orc.event_dump(5, 3000)
    total_v = total_n = 0
    macro_sum = np.zeros((len(e), 5)); dump_sum = np.zeros((3000, 5)); performed = np.zeros(3000, np.int64)
    for b in range(bands):
        monkeypatch.setenv("XSB200_BANDS", str(bands))
        monkeypatch.setenv("XSB200_BAND_INDEX", str(b))
        inp = xs.make_inputs(size="small", method="event", grid="unionized", lookups=1000, gridpoints=1000, kernel_id=6)
        sd = xs.materials_only(inp)
        with xs.move_simulation_data_to_device(inp, sd) as gpu:
            res, macro = gpu.lookup_samples(e, m, want_macro_xs=True)
            assert 0 < res.n_lookups < len(e)
            assert np.count_nonzero(macro.any(axis=1)) == res.n_lookups
            total_v += res.verification; total_n += res.n_lookups; macro_sum += macro
            de, dm, dx, da = gpu.dump(5, 3000)
            assert np.array_equal(de, oe) and np.array_equal(dm, om)            # sampling does not depend on the band
            mine = da >= 0
            assert np.array_equal(da[mine], oam[mine]) and not dx[~mine].any()
            dump_sum += dx; performed += mine
        xs.free_simulation_data(sd)
    orc.close()
    assert total_v == want_v and total_n == len(e) and np.array_equal(macro_sum, want_macro)
    assert np.all(performed == 1) and float(np.max(np.abs(dump_sum - omacro) / np.abs(omacro))) <= REL_TOL


# ---- device-side generator: byte-identical to the host generator --------------------------------------------
@pytest.mark.parametrize("size,n_gp,grid,hb", [("small", 1000, "unionized", 10000), ("small", 1000, "hash", 500),
                                               ("small", 1000, "nuclide", 10000), ("large", 300, "unionized", 10000),
                                               ("large", 300, "hash", 37), ("small", 2, "unionized", 10000),
                                               ("small", 5000, "unionized", 10000)])
def test_device_generator_byte_identical(size, n_gp, grid, hb):
    inp = xs.make_inputs(size=size, method="event", grid=grid, lookups=50000, gridpoints=n_gp, hash_bins=hb, kernel_id=4)
    host_sd = xs.grid_init_do_not_profile(inp)
    host = xs.simulation_arrays(inp, host_sd)
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        assert np.array_equal(gpu.read_array("nuclide_grid"), host["nuclide_grid"])
        if grid == "unionized":
            assert np.array_equal(gpu.read_array("unionized_energy_array"), host["unionized_energy_array"])
        if grid != "nuclide":
            assert np.array_equal(gpu.read_array("index_grid"), host["index_grid"])
        res = gpu.run()
        orc = ol.OracleProblem(inp.n_isotopes, n_gp, GRIDS[grid], hb)
        assert res.verification == orc.event(0, 50000, NTHREADS)
        orc.close()
    xs.free_simulation_data(host_sd)
    xs.free_simulation_data(mats)


# ---- radix sort: permutation, sortedness, stability ----------------------------------------------------
@pytest.mark.parametrize("n,lo,hi", [(1, 0, 32), (31, 0, 32), (4096, 0, 32), (4097, 8, 32), (1_000_003, 0, 32),
                                     (300_000, 28, 32), (300_000, 8, 32), (50_000, 0, 8)])
def test_radix_sort_properties(small, n, lo, hi):
    rng = np.random.default_rng(n + lo)
    keys = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[::7] = keys[0]                                   # plenty of duplicates
    perm = small.gpu.sort_keys(keys, lo, hi)
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32)), "must be a permutation"
    mask = np.uint32(((1 << (hi - lo)) - 1) << lo) if hi - lo < 32 else np.uint32(0xffffffff)
    digits = (keys & mask)[perm]
    assert np.all(digits[1:] >= digits[:-1]), "sorted on the selected bits"
    want = np.argsort(keys & mask, kind="stable").astype(np.uint32)
    assert np.array_equal(perm, want), "stable: equal keys keep their order"


# ---- huge lookup counts are worked through in bounded passes ------------------------------------------
def test_multi_pass_runs_match_single_pass(monkeypatch):
    monkeypatch.setenv("XSB200_MAX_PASS", "30000")
    p = Problem("small", 1000, "unionized", 500, method="event", lookups=100000)
    try:
        for k in (1, 4, 5, 6):
            inp = xs.make_inputs(size="small", grid="unionized", gridpoints=1000, hash_bins=500, method="event",
                                 lookups=100000, kernel_id=k)
            res = p.gpu.run(inp)
            assert res.verification == 302880 and res.n_lookups == 100000, k
    finally:
        p.close()


# ---- error behaviour: status codes, never exit() ------------------------------------------------------
def test_error_codes(small):
    bad_k = xs.make_inputs(size="small", gridpoints=1000, method="event", lookups=10, kernel_id=7,
                           grid={0: "unionized", 1: "nuclide", 2: "hash"}[small.inp.grid_type], hash_bins=500)
    with pytest.raises(xs.XSGpuError) as ei:            # reference: "No kernel ID" + exit(1), cuda/Main.cu:78-82
        small.gpu.run(bad_k)
    assert ei.value.code == _abi.XS_ERR_ARG
    other = xs.make_inputs(size="small", gridpoints=999, method="event", lookups=10)
    with pytest.raises(xs.XSGpuError):
        small.gpu.run(other)
    with pytest.raises(xs.XSGpuError):
        small.gpu.run_range(-1, 5)


# ---- official table at full size ------------------------------------------------------------------------
@pytest.mark.slow
def test_official_small_event_full_size():
    p = Problem("small", 11303, "unionized", 10000, method="event")
    try:
        assert p.inp.lookups == 17_000_000
        for k in (0, 6):
            inp = xs.make_inputs(size="small", method="event", kernel_id=k)
            res = p.gpu.run(inp)
            assert res.checksum == 945990 and res.n_lookups == 17_000_000
        hist = xs.make_inputs(size="small", method="history")
        assert p.gpu.run(hist).checksum == 941535
    finally:
        p.close()


@pytest.mark.slow
def test_xl_problem_golden_checksum():
    """BASELINE config 5: -s XL (355 x 238,847 grid points, 116.5 GiB unionized).  Golden value 3377 for
    -l 1000000 comes from the reference run with -G hash / -G nuclide (checksum is grid-type invariant).
    Needs ~125 GB of host RAM for the generator and ~140 GB of HBM; skipped on smaller machines."""
    import torch
    host_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    if host_gb < 150 or torch.cuda.get_device_properties(0).total_memory < 150 * 2**30:
        pytest.skip("needs 150 GB host RAM and a 180 GB GPU")
    inp = xs.make_inputs(size="XL", method="event", lookups=1000000)
    assert inp.n_gridpoints == 238847
    sd = xs.grid_init_do_not_profile(inp)
    gpu = xs.move_simulation_data_to_device(inp, sd)
    xs.free_simulation_data(sd)
    try:
        for k in (0, 4):
            res = gpu.run(xs.make_inputs(size="XL", method="event", lookups=1000000, kernel_id=k))
            assert res.checksum == 3377 and res.n_lookups == 1000000, k
    finally:
        gpu.release()


@pytest.mark.slow
def test_xxl_hash_grid_golden_checksum():
    """-s XXL (355 x 501,578 points, 8 GB nuclide grid) with -G hash, generated on the device:
    golden 344 for -l 1000000 from the reference.  XXL *unionized* (245 GiB) does not
    fit one B200 and is reported as out of memory, not silently degraded."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100 * 2**30:
        pytest.skip("needs a large GPU")
    inp = xs.make_inputs(size="XXL", method="event", grid="hash", lookups=1000000, kernel_id=4)
    assert inp.n_gridpoints == 501578
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        assert gpu.run().checksum == 344
        assert gpu.run(xs.make_inputs(size="XXL", method="event", grid="hash", lookups=1000000, kernel_id=0)).checksum == 344
    big = xs.make_inputs(size="XXL", method="event", grid="unionized", lookups=1000)
    with pytest.raises(xs.XSGpuError) as ei:
        xs.move_simulation_data_to_device(big, mats)
    assert ei.value.code == _abi.XS_ERR_CUDA
    xs.free_simulation_data(mats)


@pytest.mark.slow
def test_official_large_event_full_size():
    """The canonical FOM configuration (BASELINE.json configs[1]): 355 nuclides, 5.6 GB."""
    inp = xs.make_inputs(size="large", method="event")
    sd = xs.grid_init_do_not_profile(inp)
    gpu = xs.move_simulation_data_to_device(inp, sd)
    # the same problem built on the device must be byte-identical at full size too
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as generated:
        host = xs.simulation_arrays(inp, sd)
        assert np.array_equal(generated.read_array("unionized_energy_array"), host["unionized_energy_array"])
        assert np.array_equal(generated.read_array("nuclide_grid"), host["nuclide_grid"])
        assert np.array_equal(generated.read_array("index_grid"), host["index_grid"])
        assert generated.run(xs.make_inputs(size="large", method="event", kernel_id=4)).checksum == 952131
    xs.free_simulation_data(mats)
    xs.free_simulation_data(sd)
    try:
        for k in (0, 1, 4, 6):
            res = gpu.run(xs.make_inputs(size="large", method="event", kernel_id=k))
            assert res.checksum == 952131, k
            assert res.n_lookups == 17_000_000
        # size-independent property: any partition of the id range sums to the whole
        a = gpu.run_range(0, 6_000_000).verification + gpu.run_range(6_000_000, 11_000_000).verification
        assert a % 999983 == 952131
        assert gpu.run(xs.make_inputs(size="large", method="history")).checksum == 954318
    finally:
        gpu.release()


@pytest.mark.gpu
@pytest.mark.slow
def test_billion_lookups_bounded_passes():
    """BASELINE config 5: 10^9 lookups worked through in bounded passes (XSB200_MAX_PASS, 2^26 by
    default) on the device-generated large problem.  Golden: the reference CUDA build's output
    (260078) corrected for the int overflow of its thrust::reduce(..., 0) at sums above 2^31
    (cuda/Simulation.cu:34): (260078 + 35965) mod 999983 = 296043."""
    inp = xs.make_inputs(size="large", method="event", lookups=1_000_000_000, kernel_id=6)
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        res = gpu.run(inp)
        assert res.n_lookups == 1_000_000_000 and res.checksum == 296043
        assert xs.expected_checksum(inp) == 296043
        assert res.verification > 2**31
        k4 = gpu.run(xs.make_inputs(size="large", method="event", lookups=1_000_000_000, kernel_id=4))
        assert k4.verification == res.verification
        # any partition of the id range sums to the whole
        a = gpu.run_range(0, 400_000_000, inp).verification + gpu.run_range(400_000_000, 600_000_000, inp).verification
        assert a == res.verification
    xs.free_simulation_data(mats)


# ---- two live contexts on one device keep their own material tables ------------------------------------
def test_two_live_contexts_do_not_clobber_each_other():
    """Round 1 kept the zero-padded concentrations in a process-global __constant__ symbol that every
    xs_gpu_init overwrote: with a 68-nuclide and a 355-nuclide context alive together the older one
    silently multiplied by the newer one's concentrations on the -k 4/5/6 / host-sample paths.  The
    table is a kernel parameter now; both contexts must keep giving their single-context results
    whatever the order of creation and use."""
    rng = np.random.default_rng(23)
    e = rng.random(60_000); m = rng.integers(0, 12, len(e)).astype(np.int32)
    a = Problem("small", 1000, "unionized", 500, method="event", lookups=100000)
    want_a = a.oracle.lookup_samples(e, m)
    b = Problem("large", 1000, "unionized", 500, method="event", lookups=100000)     # created while `a` is alive
    want_b = b.oracle.lookup_samples(e, m)
    try:
        for _ in range(2):
            for p, (want_v, want_macro), golden, size in ((a, want_a, 302880, "small"), (b, want_b, 303045, "large")):
                for k in (4, 5, 6):
                    inp = xs.make_inputs(size=size, grid="unionized", gridpoints=1000, hash_bins=500, method="event",
                                         lookups=100000, kernel_id=k)
                    assert p.gpu.run(inp).verification == golden, (size, k)
                res, macro = p.gpu.lookup_samples(e, m, want_macro_xs=True)
                assert res.verification == want_v and np.array_equal(macro, want_macro), size
        # and a third context of the first size, created last, does not disturb the second
        c = Problem("small", 1000, "hash", 500, method="event", lookups=100000)
        try:
            res, macro = b.gpu.lookup_samples(e, m, want_macro_xs=True)
            assert res.verification == want_b[0] and np.array_equal(macro, want_b[1])
            res, macro = c.gpu.lookup_samples(e, m, want_macro_xs=True)
            assert res.verification == want_a[0] and np.array_equal(macro, want_a[1])
        finally:
            c.close()
    finally:
        a.close(); b.close()


# ---- host samples are validated, not trusted ---------------------------------------------------------------
@pytest.mark.parametrize("bad_e,bad_m", [(0.5, 12), (0.5, -1), (0.5, 1 << 20), (-1e-9, 3), (1.0000001, 3), (float("nan"), 3),
                                          (float("inf"), 0), (-float("inf"), 11)])
def test_lookup_samples_rejects_bad_samples(small, bad_e, bad_m):
    """A material outside [0, 12) or an energy outside [0, 1] is XS_ERR_ARG (include/xs_gpu.h), not an
    out-of-bounds access; the context stays usable afterwards."""
    rng = np.random.default_rng(1)
    e = rng.random(5000); m = rng.integers(0, 12, len(e)).astype(np.int32)
    good, _ = small.gpu.lookup_samples(e, m)
    e2, m2 = e.copy(), m.copy()
    e2[1234], m2[1234] = bad_e, bad_m
    with pytest.raises(xs.XSGpuError) as ei:
        small.gpu.lookup_samples(e2, m2, want_macro_xs=True)
    assert ei.value.code == _abi.XS_ERR_ARG
    again, _ = small.gpu.lookup_samples(e, m)
    assert again.verification == good.verification and again.n_lookups == len(e)
    edge, _ = small.gpu.lookup_samples(np.array([0.0, 1.0]), np.array([0, 11], np.int32))     # the closed ends are fine
    assert edge.n_lookups == 2


# ---- per-lookup parity at the canonical grid size (355 x 11,303) ---------------------------------------
@pytest.fixture(scope="module")
def canonical_samples():
    """The fuel lookups (all 2.36 M: 209 per grid interval, the density xs_dense_kernel is built for)
    and every fourth of the others of the canonical 17 M-lookup stream, with the oracle's macro_xs
    for them.  The oracle works on the hash grid (200 MB; macro_xs is grid-type invariant bit for
    bit: tests/test_oracle.py::test_macro_xs_bit_identical_to_reference)."""
    n = 17_000_000
    e = np.empty(n); m = np.empty(n, np.int32)
    ol.oracle().xo_sample(0, n, e.ctypes.data, m.ctypes.data)
    keep = (m == 0) | (np.arange(n) % 4 == 0)
    e, m = e[keep].copy(), m[keep].copy()
    orc = ol.OracleProblem(355, 11303, 2, 10000)
    v, macro = orc.lookup_samples(e, m, nthreads=NTHREADS)
    orc.close()
    return e, m, v, macro


@pytest.mark.slow
@pytest.mark.parametrize("grid", ["unionized", "hash", "nuclide"])
def test_canonical_grid_per_lookup_parity(canonical_samples, grid):
    """-k 6 pipeline (xs_gpu_lookup_samples: sort + xs_dense_kernel for fuel and the other dense
    materials + xs_sorted_kernel for the rest) on the REAL problem size: 355 nuclides x 11,303 grid
    points, fuel at its real density (11-chunk nuclide loop, index-row prefetch of the chunk after
    next, 4-record rings).  6 M lookups, every macro_xs vector compared bit for bit with the oracle,
    plus the committed reference rows of this grid size."""
    e, m, v, omacro = canonical_samples
    inp = xs.make_inputs(size="large", method="event", grid=grid, kernel_id=6)
    assert (inp.n_isotopes, inp.n_gridpoints) == (355, 11303)
    mats = xs.materials_only(inp)                       # device-side generator: byte-identical to the host's
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        assert np.count_nonzero(m == 0) >= 64 * 11303   # fuel is dense at this count
        res, macro = gpu.lookup_samples(e, m, want_macro_xs=True)
        assert res.n_lookups == len(e) and res.verification == v
        assert np.array_equal(macro, omacro)
        rows = [g for g in GOLDEN["lookups"] if (g["n_isotopes"], g["n_gridpoints"]) == (355, 11303)]
        assert rows
        for row in rows[0]["rows"]:
            ge, gm, gx, ga = gpu.dump(row["id"], 1)
            assert ge[0] == float.fromhex(row["energy"]) and gm[0] == row["mat"] and ga[0] == row["argmax"]
            assert max_rel(gx[0], unhex(row["macro_xs"])) <= REL_TOL
            r1, x1 = gpu.lookup_samples(ge, gm, want_macro_xs=True)          # sweep path: reference order
            assert np.array_equal(x1[0], unhex(row["macro_xs"]))
        assert gpu.run_range(0, 100000, xs.make_inputs(size="large", method="event", grid=grid, kernel_id=6, lookups=100000)).verification == 298914
    xs.free_simulation_data(mats)


def test_small_canonical_grid_golden_rows():
    """The committed reference rows for 68 x 11,303 (tests/golden/reference_vectors.json) on the GPU."""
    rows = [g for g in GOLDEN["lookups"] if (g["n_isotopes"], g["n_gridpoints"]) == (68, 11303)]
    assert rows
    inp = xs.make_inputs(size="small", method="event", grid="unionized", kernel_id=6)
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats) as gpu:
        for row in rows[0]["rows"]:
            ge, gm, gx, ga = gpu.dump(row["id"], 1)
            assert ge[0] == float.fromhex(row["energy"]) and gm[0] == row["mat"] and ga[0] == row["argmax"]
            assert max_rel(gx[0], unhex(row["macro_xs"])) <= REL_TOL
            _, x1 = gpu.lookup_samples(ge, gm, want_macro_xs=True)
            assert np.array_equal(x1[0], unhex(row["macro_xs"]))
    xs.free_simulation_data(mats)


# ---- the host driver binary: exit status = checksum validity (the reference's only test) ---------------
XSBENCH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xsbench_b200", "xsbench")


def run_xsbench(args, cwd=None):
    import subprocess
    p = subprocess.run([XSBENCH] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return p.returncode, p.stdout


@pytest.mark.parametrize("args,checksum", [
    (["-s", "small", "-m", "event", "-k", "0"], 945990),
    (["-s", "small", "-m", "event", "-k", "6"], 945990),
    (["-s", "small", "-m", "event", "-k", "4", "-G", "hash"], 945990),
    (["-s", "small", "-m", "event", "-k", "6", "-G", "nuclide"], 945990),
    (["-s", "small", "-m", "history"], 941535),
    (["-s", "small", "-m", "event", "-k", "6", "--device-init"], 945990),
])
def test_xsbench_binary_exit_status(args, checksum):
    """openmp-threading/Main.c:112,122 + .github/workflows/omp.yml:20-27: the reference is tested by
    running the binary on a table configuration and checking its exit status.  Same here, on the GPU."""
    rc, out = run_xsbench(args + ["-t", str(NTHREADS)])
    assert rc == 0, out[-2000:]
    assert f"Verification checksum: {checksum} (Valid)" in out


def test_xsbench_binary_reports_invalid_and_errors():
    rc, out = run_xsbench(["-s", "small", "-m", "event", "-l", "12345", "-g", "300"])     # not in the table
    assert rc == 1 and "WARNING - INVALID CHECKSUM" in out
    rc, out = run_xsbench(["-s", "small", "-m", "event", "-k", "9", "-g", "300"])
    assert rc != 0 and "No kernel ID 9" in out
    rc, out = run_xsbench(["-s", "nonsense"])
    assert rc == 4                                                                       # usage error (cuda/io.cu:208-224)


@pytest.mark.parametrize("grid", ["unionized", "hash"])
def test_xsbench_binary_write_then_read(tmp_path, grid):
    """-b write, then -b read in the same directory (cuda/io.cu:443-495): the problem loaded from
    XS_data.dat gives the same (valid) checksum as the generated one; so does a file written from a
    problem that was generated on the device."""
    base = ["-s", "small", "-m", "event", "-G", grid, "-k", "6", "-l", "100000", "-t", str(NTHREADS)]
    rc, out = run_xsbench(base + ["-b", "write"], cwd=tmp_path)
    assert rc == 0 and "Verification checksum: 299541 (Valid)" in out, out[-1500:]
    first = open(tmp_path / "XS_data.dat", "rb").read()
    rc, out = run_xsbench(base + ["-b", "read"], cwd=tmp_path)
    assert rc == 0 and "Reading all data structures from binary file" in out
    assert "Verification checksum: 299541 (Valid)" in out
    rc, out = run_xsbench(base + ["-b", "write", "--device-init"], cwd=tmp_path)
    assert rc == 0, out[-1500:]
    assert open(tmp_path / "XS_data.dat", "rb").read() == first                         # byte-identical generator
    rc, out = run_xsbench(["-s", "small", "-m", "history", "-G", grid, "-p", "3000", "-b", "read"], cwd=tmp_path)
    assert "Verification checksum:" in out and rc in (0, 1)


# ---- the fused-arithmetic build of the dense kernel (XSB200_ARITH=fused, not the default) ---------------------
@pytest.mark.parametrize("grid,hb", [("unionized", 50), ("hash", 50), ("nuclide", 50)])
def test_fused_arithmetic_within_contract(monkeypatch, grid, hb):
    """12 instead of 24 FP64 operations per (lookup, nuclide): macro_xs within the 1e-12 contract (observed: a few
    ulp), every integer (argmax checksum) bit-exact -- a lookup whose two largest channels are within 1e-10 is
    recomputed in the reference's order.  Same configuration as test_dense_kernel_at_the_default_split: every
    material goes through xs_dense_kernel."""
    monkeypatch.setenv("XSB200_ARITH", "fused")
    p = Problem("small", 100, grid, hb, method="event", lookups=1_000_000)
    try:
        assert p.gpu.info().fp64_ops_per_pair == 12
        inp = xs.make_inputs(size="small", grid=grid, gridpoints=100, hash_bins=hb, method="event", lookups=1_000_000, kernel_id=6)
        assert p.gpu.run(inp).verification == p.oracle.event(0, 1_000_000, NTHREADS)
        rng = np.random.default_rng(19)
        g = p.oracle.nuclide_grid[0::6]
        e = np.concatenate([rng.random(300_000), g, np.nextafter(g, 0), np.nextafter(g, 1)])
        e = e[(e >= 0) & (e < 1)]
        m = rng.integers(0, 12, len(e)).astype(np.int32)
        res, macro = p.gpu.lookup_samples(e, m, want_macro_xs=True)
        v, omacro = p.oracle.lookup_samples(e, m)
        assert res.verification == v and res.n_lookups == len(e)
        assert max_rel(macro, omacro) <= REL_TOL
        assert not np.array_equal(macro, omacro)            # (it really is the other arithmetic)
    finally:
        p.close()
    monkeypatch.delenv("XSB200_ARITH")
    q = Problem("small", 100, grid, hb, method="event", lookups=1000)
    try:
        assert q.gpu.info().fp64_ops_per_pair == 24
    finally:
        q.close()


# ---- the driver's smoke() entry point itself -----------------------------------------------------------------------
def test_graft_entry_smoke():
    """__graft_entry__.smoke() is what the driver runs on the GPU box before the bench: keep it in the suite so a
    change in the pipeline (launch counts, defaults) cannot break it unnoticed."""
    import __graft_entry__ as entry
    entry.smoke()
