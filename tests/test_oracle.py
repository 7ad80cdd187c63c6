"""Pin the CPU oracle (oracle/xs_oracle.c) before anything is checked against it.

Three independent anchors:
  1. the reference's own 4-entry checksum table (openmp-threading/io.c:85-96);
  2. committed golden vectors generated from the unmodified reference
     (tests/golden/reference_vectors.json, made by tests/golden/make_golden.py);
  3. when oracle/_ref/libxsref.so is present (build container, GPU box): the reference's own
     functions, called directly -- byte-identical data, bit-identical macro_xs.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
NTHREADS = os.cpu_count() or 1


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


# ---- 1. official table ----------------------------------------------------------------------
def test_official_event_small():
    p = ol.OracleProblem(68, 11303, 0)
    assert p.event(0, 17_000_000, NTHREADS) % 999983 == GOLDEN["official_table"]["event_small"] == 945990


def test_official_history_small():
    p = ol.OracleProblem(68, 11303, 2)          # checksum is grid-type invariant (README.md:164)
    assert p.history(0, 500_000, 34, NTHREADS) % 999983 == GOLDEN["official_table"]["history_small"] == 941535


@pytest.mark.slow
def test_official_event_large():
    p = ol.OracleProblem(355, 11303, 2)         # hash grid: same checksum, 200 MB instead of 5.6 GB
    assert p.event(0, 17_000_000, NTHREADS) % 999983 == GOLDEN["official_table"]["event_large"] == 952131


@pytest.mark.slow
def test_official_history_large():
    p = ol.OracleProblem(355, 11303, 2)
    assert p.history(0, 500_000, 34, NTHREADS) % 999983 == GOLDEN["official_table"]["history_large"] == 954318


# ---- 2. committed golden vectors ------------------------------------------------------------
@pytest.mark.parametrize("g", GOLDEN["checksums"], ids=lambda g: "iso{n_isotopes}-gp{n_gridpoints}-G{grid_type}-m{method}-l{lookups}-p{particles}".format(**g))
def test_golden_checksums(g):
    p = ol.OracleProblem(g["n_isotopes"], g["n_gridpoints"], g["grid_type"], g["hash_bins"])
    if g["method"] == 2:
        v = p.event(0, g["lookups"], NTHREADS)
    else:
        v = p.history(0, g["particles"], g["lookups"], NTHREADS)
    assert v == g["verification"]


@pytest.mark.parametrize("g", GOLDEN["lookups"], ids=lambda g: "iso{n_isotopes}-gp{n_gridpoints}-G{grid_type}".format(**g))
def test_golden_macro_xs_bit_exact(g):
    p = ol.OracleProblem(g["n_isotopes"], g["n_gridpoints"], g["grid_type"], g["hash_bins"])
    for row in g["rows"]:
        _, e, m, x, a = p.event_dump(row["id"], 1)
        assert e[0] == float.fromhex(row["energy"])
        assert m[0] == row["mat"]
        assert np.array_equal(x[0], unhex(row["macro_xs"]))       # same order of operations => same bits
        assert a[0] == row["argmax"]


def test_golden_generator_fingerprints():
    for f in GOLDEN["generator_fingerprints"]:
        p = ol.OracleProblem(f["n_isotopes"], f["n_gridpoints"], f["grid_type"], f["hash_bins"])
        assert p.num_nucs.tolist() == f["num_nucs"] and p.max_num_nucs == f["max_num_nucs"]
        assert np.array_equal(p.nuclide_grid[:12], unhex(f["nuclide_grid_head"]))
        assert np.array_equal(p.nuclide_grid[-12:], unhex(f["nuclide_grid_tail"]))
        assert float(np.sum(p.nuclide_grid)) == float.fromhex(f["nuclide_grid_sum"])
        assert int(np.sum(p.index_grid.astype(np.int64))) == f["index_grid_sum"]
        ig = p.index_grid.astype(np.int64)
        assert int(np.bitwise_xor.reduce(ig * (np.arange(len(ig)) % 1000003 + 1))) == f["index_grid_crc"]
        w = p.max_num_nucs
        valid = np.concatenate([np.arange(m * w, m * w + p.num_nucs[m]) for m in range(12)])
        assert float(np.sum(p.concs[valid])) == float.fromhex(f["concs_valid_sum"])
        assert int(np.sum(p.mats[valid])) == f["mats_valid_sum"]
        if f["grid_type"] == 0:
            assert np.array_equal(p.ueg[:4], unhex(f["ueg_head"]))
            assert np.array_equal(p.ueg[-4:], unhex(f["ueg_tail"]))


# ---- 3. against the reference's own functions -----------------------------------------------
needs_ref = pytest.mark.skipif(not ol.have_reference(), reason="oracle/_ref/libxsref.so not built")


@needs_ref
def test_lcg_and_pick_mat_match_reference():
    r, o = ol.reference(), ol.oracle()
    rng = np.random.default_rng(1)
    for n in [0, 1, 2, 3, 1000, 2**31, 2**40 + 12345, 2**62 + 7] + rng.integers(0, 2**62, 50).tolist():
        for seed in (1070, 42, 1070 * 1070, 2**63 - 1):
            assert r.fast_forward_LCG(seed, n) == o.xo_lcg_skip(seed, n)
    a = C.c_uint64(1070); b = C.c_uint64(1070)
    for _ in range(5000):
        assert r.LCG_random_double(C.byref(a)) == o.xo_lcg_next(C.byref(b))
        assert a.value == b.value
    a = C.c_uint64(99); b = C.c_uint64(99)
    picks = []
    for _ in range(20000):
        x, y = r.pick_mat(C.byref(a)), o.xo_pick_mat(C.byref(b))
        assert x == y and a.value == b.value
        picks.append(x)
    assert set(picks) == set(range(12))


@needs_ref
@pytest.mark.parametrize("n_iso,n_gp,gt,hb", [(68, 1000, 0, 10000), (68, 1000, 2, 500), (68, 1000, 1, 10000),
                                              (355, 300, 0, 10000), (355, 300, 2, 37), (68, 2, 0, 10000),
                                              (68, 3, 2, 1)])
def test_generated_data_byte_identical_to_reference(n_iso, n_gp, gt, hb):
    r = ol.reference()
    inp = ol.ref_inputs(n_iso, n_gp, gt, hb)
    sd = r.grid_init_do_not_profile(inp, 1)
    p = ol.OracleProblem(n_iso, n_gp, gt, hb)
    assert np.array_equal(np.ctypeslib.as_array(sd.nuclide_grid, shape=(n_iso * n_gp * 6,)), p.nuclide_grid)
    assert sd.length_index_grid == len(p.index_grid)
    if len(p.index_grid):
        assert np.array_equal(np.ctypeslib.as_array(sd.index_grid, shape=(len(p.index_grid),)), p.index_grid)
    if gt == 0:
        assert np.array_equal(np.ctypeslib.as_array(sd.unionized_energy_array, shape=(n_iso * n_gp,)), p.ueg)
    w = sd.max_num_nucs
    assert w == p.max_num_nucs
    nn = np.ctypeslib.as_array(sd.num_nucs, shape=(12,))
    assert np.array_equal(nn, p.num_nucs)
    mats = np.ctypeslib.as_array(sd.mats, shape=(12 * w,))
    concs = np.ctypeslib.as_array(sd.concs, shape=(12 * w,))
    for m in range(12):                      # padding is uninitialised in the reference
        assert np.array_equal(mats[m * w:m * w + nn[m]], p.mats[m * w:m * w + nn[m]])
        assert np.array_equal(concs[m * w:m * w + nn[m]], p.concs[m * w:m * w + nn[m]])


@needs_ref
@pytest.mark.parametrize("gt,hb", [(0, 10000), (2, 300), (1, 10000)])
def test_macro_xs_bit_identical_to_reference(gt, hb):
    r = ol.reference()
    n_iso, n_gp = 68, 1500
    inp = ol.ref_inputs(n_iso, n_gp, gt, hb)
    sd = r.grid_init_do_not_profile(inp, 1)
    p = ol.OracleProblem(n_iso, n_gp, gt, hb)
    rng = np.random.default_rng(7)
    energies = np.concatenate([rng.random(300), [0.0, 1e-300, 1e-9, 0.5, 1.0 - 2**-53, p.nuclide_grid[0], p.nuclide_grid[6]]])
    mats = rng.integers(0, 12, len(energies)).astype(np.int32)
    _, x = p.lookup_samples(energies, mats)
    for i, (e, m) in enumerate(zip(energies, mats)):
        assert np.array_equal(ol.ref_macro_xs(inp, sd, float(e), int(m)), x[i]), (i, e, m)


@needs_ref
def test_ueg_search_matches_reference_and_closed_form():
    r, o = ol.reference(), ol.oracle()
    p = ol.OracleProblem(68, 400, 0)
    n = len(p.ueg)
    rng = np.random.default_rng(3)
    qs = np.concatenate([rng.random(2000), p.ueg[:5], p.ueg[-5:], p.ueg[rng.integers(0, n, 200)], [0.0, 1.0, -1.0, 2.0]])
    ub = np.searchsorted(p.ueg, qs, side="right")
    closed = np.clip(ub - 1, 0, n - 2)
    for q, c in zip(qs, closed):
        a = r.grid_search(n, float(q), p.ueg.ctypes.data)
        b = o.xo_search_ueg(n, float(q), p.ueg.ctypes.data)
        assert a == b == c


@needs_ref
def test_event_and_history_match_reference_drivers():
    r = ol.reference()
    for gt, hb in ((0, 10000), (2, 200)):
        inp = ol.ref_inputs(68, 800, gt, hb, lookups=30000)
        sd = r.grid_init_do_not_profile(inp, 1)
        p = ol.OracleProblem(68, 800, gt, hb)
        assert r.run_event_based_simulation(inp, sd, 1) == p.event(0, 30000, NTHREADS)
        inp_h = ol.ref_inputs(68, 800, gt, hb, lookups=13, particles=2000, method=1)
        assert r.run_history_based_simulation(inp_h, sd, 1) == p.history(0, 2000, 13, NTHREADS)


@needs_ref
@pytest.mark.parametrize("n_iso,n_gp,gt,hb,lookups", [(68, 800, 0, 10000, 30000), (68, 800, 2, 200, 30000),
                                                        (68, 800, 1, 10000, 20000), (355, 300, 0, 10000, 25000)])
def test_reference_sorted_variant_matches_oracle(n_iso, n_gp, gt, hb, lookups):
    """the reference's CPU "-k 1" (sample, key-value quicksort by material, per-material
    quicksort by energy, 12 sorted loops: openmp-threading/Simulation.c:698-871 with the sorts of :565-679)
    run unmodified from libxsref.so.  The verification sum does not depend on the order of the lookups, so
    it must equal the reference's own unsorted driver and the oracle -- which is what pins the hash of the
    GPU's sorted pipeline (-k 6) to the reference's sorted path, not only to its baseline."""
    r = ol.reference()
    inp = ol.ref_inputs(n_iso, n_gp, gt, hb, lookups=lookups, hm=b"small" if n_iso == 68 else b"large", nthreads=NTHREADS)
    sd = r.grid_init_do_not_profile(inp, 1)
    k0 = r.run_event_based_simulation(inp, sd, 1)
    k1 = r.run_event_based_simulation_optimization_1(inp, sd, 1)
    p = ol.OracleProblem(n_iso, n_gp, gt, hb)
    assert k1 == k0 == p.event(0, lookups, NTHREADS)
    p.close()


def test_event_partition_is_exact():
    p = ol.OracleProblem(68, 300, 0)
    whole = p.event(0, 10000, NTHREADS)
    assert whole == p.event(0, 3333, 1) + p.event(3333, 4000, 2) + p.event(7333, 2667, NTHREADS)
    v, e, m, x, a = p.event_dump(0, 10000)
    assert v == whole and v == int(np.sum(a + 1))
    e2 = np.empty(10000); m2 = np.empty(10000, np.int32)
    ol.oracle().xo_sample(0, 10000, e2.ctypes.data, m2.ctypes.data)
    assert np.array_equal(e, e2) and np.array_equal(m, m2)
    v2, x2 = p.lookup_samples(e, m)
    assert v2 == whole and np.array_equal(x, x2)
