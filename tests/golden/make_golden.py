"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libxsref.so,
compiled from /root/reference/openmp-threading by `make -C oracle ref`).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
The JSON files are committed; tests read them, never the reference.
Floating-point values are stored as C99 hex strings so they round-trip bit-exactly.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def hexlist(a):
    return [float(x).hex() for x in np.asarray(a).ravel()]


def per_lookup_vectors(n_iso, n_gp, grid_type, hash_bins, ids):
    """(E, mat, macro_xs[5], argmax) for event-mode lookup ids, from the reference's own
    fast_forward_LCG / LCG_random_double / pick_mat / calculate_macro_xs."""
    r = ol.reference()
    inp = ol.ref_inputs(n_iso, n_gp, grid_type, hash_bins)
    sd = r.grid_init_do_not_profile(inp, 1)
    rows = []
    for i in ids:
        seed = C.c_uint64(r.fast_forward_LCG(1070, 2 * i))
        e = r.LCG_random_double(C.byref(seed))
        mat = r.pick_mat(C.byref(seed))
        macro = ol.ref_macro_xs(inp, sd, e, mat)
        rows.append({"id": int(i), "energy": float(e).hex(), "mat": int(mat), "macro_xs": hexlist(macro),
                     "argmax": int(np.argmax(macro))})
    return rows


def checksum(n_iso, n_gp, grid_type, hash_bins, method, lookups, particles, hm=b"small"):
    r = ol.reference()
    inp = ol.ref_inputs(n_iso, n_gp, grid_type, hash_bins, lookups, particles, method, hm)
    sd = r.grid_init_do_not_profile(inp, 1)
    if method == 2:
        v = r.run_event_based_simulation(inp, sd, 1)
    else:
        v = r.run_history_based_simulation(inp, sd, 1)
    return int(v)


def main():
    if not ol.have_reference():
        sys.exit("oracle/_ref/libxsref.so missing: run `make -C oracle ref` first")
    out = {"generator": "tests/golden/make_golden.py", "reference": "ANL-CESAR/XSBench v20 openmp-threading",
           "official_table": {"event_small": 945990, "event_large": 952131,
                              "history_small": 941535, "history_large": 954318},
           "checksums": [], "lookups": []}

    # un-modded verification sums for small, fast configurations (grid-type invariant)
    for (n_iso, n_gp, gt, hb, method, lookups, particles) in [
        (68, 1000, 0, 10000, 2, 100000, 0),
        (68, 1000, 2, 500, 2, 100000, 0),
        (68, 1000, 1, 10000, 2, 100000, 0),
        (68, 500, 0, 10000, 2, 20000, 0),
        (68, 11303, 0, 10000, 2, 100000, 0),
        (68, 11303, 2, 10000, 2, 1000000, 0),
        (355, 1000, 0, 10000, 2, 100000, 0),
        (355, 1000, 2, 2000, 2, 100000, 0),
        (68, 1000, 0, 10000, 1, 34, 3000),
        (68, 11303, 2, 10000, 1, 34, 10000),
        (355, 1000, 2, 2000, 1, 34, 3000),
        (355, 1000, 0, 10000, 1, 7, 5000),
        (355, 11303, 2, 10000, 2, 100000, 0),      # the canonical grid size (hash grid: same sums, fast to build)
    ]:
        v = checksum(n_iso, n_gp, gt, hb, method, lookups, particles)
        out["checksums"].append({"n_isotopes": n_iso, "n_gridpoints": n_gp, "grid_type": gt, "hash_bins": hb,
                                 "method": method, "lookups": lookups, "particles": particles,
                                 "verification": v, "checksum": v % 999983})
        print(out["checksums"][-1])

    # per-lookup macro_xs vectors
    ids = list(range(0, 64)) + [1000, 4097, 99999, 1234567, 16999999, 2**30 - 1]
    for (n_iso, n_gp, gt, hb) in [(68, 1000, 0, 10000), (68, 1000, 2, 500), (68, 1000, 1, 10000),
                                  (355, 1000, 0, 10000), (68, 11303, 0, 10000), (355, 11303, 2, 10000)]:
        out["lookups"].append({"n_isotopes": n_iso, "n_gridpoints": n_gp, "grid_type": gt, "hash_bins": hb,
                               "rows": per_lookup_vectors(n_iso, n_gp, gt, hb, ids)})
        print("vectors", n_iso, n_gp, gt)

    # generator fingerprints: a few entries of each generated array
    r = ol.reference()
    finger = []
    for (n_iso, n_gp, gt, hb) in [(68, 1000, 0, 10000), (68, 1000, 2, 500), (355, 200, 0, 10000)]:
        inp = ol.ref_inputs(n_iso, n_gp, gt, hb)
        sd = r.grid_init_do_not_profile(inp, 1)
        ng = np.ctypeslib.as_array(sd.nuclide_grid, shape=(n_iso * n_gp * 6,))
        ig = np.ctypeslib.as_array(sd.index_grid, shape=(sd.length_index_grid,))
        concs = np.ctypeslib.as_array(sd.concs, shape=(sd.length_concs,))
        mats = np.ctypeslib.as_array(sd.mats, shape=(sd.length_mats,))
        nn = np.ctypeslib.as_array(sd.num_nucs, shape=(12,))
        w = sd.max_num_nucs
        valid = np.concatenate([np.arange(m * w, m * w + nn[m]) for m in range(12)])
        entry = {"n_isotopes": n_iso, "n_gridpoints": n_gp, "grid_type": gt, "hash_bins": hb,
                 "num_nucs": nn.tolist(), "max_num_nucs": int(w),
                 "nuclide_grid_head": hexlist(ng[:12]), "nuclide_grid_tail": hexlist(ng[-12:]),
                 "nuclide_grid_sum": float(np.sum(ng)).hex(),
                 "index_grid_sum": int(np.sum(ig.astype(np.int64))),
                 "index_grid_crc": int(np.bitwise_xor.reduce(ig.astype(np.int64) * (np.arange(len(ig)) % 1000003 + 1))),
                 "concs_valid_sum": float(np.sum(concs[valid])).hex(),
                 "mats_valid_sum": int(np.sum(mats[valid]))}
        if gt == 0:
            ueg = np.ctypeslib.as_array(sd.unionized_energy_array, shape=(sd.length_unionized_energy_array,))
            entry["ueg_head"] = hexlist(ueg[:4]); entry["ueg_tail"] = hexlist(ueg[-4:])
            entry["ueg_sum"] = float(np.sum(ueg)).hex()
        finger.append(entry)
    out["generator_fingerprints"] = finger

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote reference_vectors.json")


if __name__ == "__main__":
    main()
