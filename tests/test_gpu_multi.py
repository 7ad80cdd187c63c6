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
