"""Multi-GPU paths (need >= 2 GPUs on the box; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol
import xsbench_b200 as xs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_in_process_two_gpus_nccl_allreduce():
    """xs_gpu_init(n_gpus=2): grid replicated by peer copy, ids split, NCCL all-reduce of the result."""
    inp = xs.make_inputs(size="small", method="event", grid="unionized", lookups=100000, gridpoints=1000)
    sd = xs.grid_init_do_not_profile(inp)
    with xs.move_simulation_data_to_device(inp, sd, n_gpus=2) as gpu:
        for k in (0, 4, 6):
            res = gpu.run(xs.make_inputs(size="small", method="event", grid="unionized", lookups=100000,
                                         gridpoints=1000, kernel_id=k))
            assert res.n_gpus == 2 and res.n_lookups == 100000 and res.verification == 302880
    xs.free_simulation_data(sd)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_weak_scaling():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--size", "small", "--lookups", "1000000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["lookups_counted"] == 2_000_000 and line["scaling"] == "weak"
    want = ol.OracleProblem(68, 11303, 2).event(0, 2_000_000, os.cpu_count() or 1) % 999983
    assert line["checksum"] == want


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_in_process_energy_bands(monkeypatch):
    """XSB200_BANDS=2 on 2 GPUs: each GPU holds half of the index rows, draws all ids, keeps its band;
    the NCCL all-reduce adds the two halves up."""
    monkeypatch.setenv("XSB200_BANDS", "2")
    inp = xs.make_inputs(size="small", method="event", grid="unionized", lookups=100000, gridpoints=1000, kernel_id=6)
    for sd in (xs.grid_init_do_not_profile(inp), xs.materials_only(inp)):
        with xs.move_simulation_data_to_device(inp, sd, n_gpus=2) as gpu:
            res = gpu.run(inp)
            assert res.n_gpus == 2 and res.n_lookups == 100000 and res.verification == 302880
        xs.free_simulation_data(sd)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_in_process_history_and_host_samples_on_several_gpus(monkeypatch):
    """History mode and xs_gpu_lookup_samples with one process driving 2 GPUs: particles / samples split (grid
    replicated), and -- XSB200_BANDS=2 -- every GPU stepping every particle, looking up its energy band and the
    feedback bytes all-reduced (NCCL, uint8) every generation."""
    import numpy as np
    orc = ol.OracleProblem(68, 1000, 0, 10000)
    want_hist = orc.history(0, 3000, 34, os.cpu_count() or 1)
    rng = np.random.default_rng(31)
    e = rng.random(50_000); m = rng.integers(0, 12, len(e)).astype(np.int32)
    want_v, want_macro = orc.lookup_samples(e, m)
    orc.close()
    hist = xs.make_inputs(size="small", method="history", grid="unionized", gridpoints=1000, particles=3000, lookups=34)
    for bands in (None, "2"):
        if bands:
            monkeypatch.setenv("XSB200_BANDS", bands)
        sd = xs.materials_only(hist)
        with xs.move_simulation_data_to_device(hist, sd, n_gpus=2) as gpu:
            res = gpu.run(hist)
            assert res.n_gpus == 2 and res.n_lookups == 34 * 3000 and res.verification == want_hist == 309181
            r2, macro = gpu.lookup_samples(e, m, want_macro_xs=True)
            assert r2.verification == want_v and r2.n_lookups == len(e) and np.array_equal(macro, want_macro)
        xs.free_simulation_data(sd)


@pytest.mark.slow
@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_xxl_unionized_across_gpus():
    """XXL unionized (253 GB of index rows) does not fit one B200: with several GPUs the grid is
    sharded by energy band automatically.  Golden 344 (the hash is grid-type invariant)."""
    inp = xs.make_inputs(size="XXL", method="event", grid="unionized", lookups=1_000_000, kernel_id=6)
    mats = xs.materials_only(inp)
    with xs.move_simulation_data_to_device(inp, mats, n_gpus=_n_gpus()) as gpu:
        res = gpu.run(inp)
        assert res.n_lookups == 1_000_000 and res.checksum == 344
        # default lookup count: 33803, pinned by the CPU oracle on the XXL nuclide grid (the hash does
        # not depend on the grid type)
        res = gpu.run(xs.make_inputs(size="XXL", method="event", grid="unionized", kernel_id=6))
        assert res.n_lookups == 17_000_000 and res.checksum == 33803
    xs.free_simulation_data(mats)
