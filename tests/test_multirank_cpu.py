"""world_size-2 checks of the N>1 host logic on CPU (gloo): id sharding + the all-reduce of
{verification, n_lookups}.  The per-rank work is done by the oracle here (no GPU in this
container); on the GPU box the same sharding drives xs_gpu_run_range (tests/test_gpu_multi.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, total, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from xsbench_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ol.OracleProblem(68, 300, 0)
    if mode == "strong":
        first, count = sharding.strong_shard(total, rank, world)
    else:
        first, count = sharding.weak_shard(total, rank, world)
    v = p.event(first, count, 1)
    tv, tn = sharding.allreduce_result(v, count)
    # bench.py's per-step collective (asynchronous submit, one waiting read), twice over like two steps
    red = sharding.ResultReducer()
    red.submit(1, 2)
    red.submit(v, count)
    assert red.result() == (tv, tn)
    q.put((rank, first, count, tv, tn))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,total", [("strong", 9001), ("weak", 4000)])
def test_two_rank_sharding_and_allreduce(mode, total):
    import oracle_lib as ol
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = total if mode == "strong" else 2 * total
    want = ol.OracleProblem(68, 300, 0).event(0, whole, 2)
    assert out[0][1] == 0 and out[0][1] + out[0][2] == out[1][1]          # contiguous, no gap
    assert out[1][1] + out[1][2] == whole
    for r in out:
        assert r[3] == want and r[4] == whole                              # same on every rank


def test_shard_helpers():
    from xsbench_b200 import sharding
    for total in (0, 1, 7, 17_000_000):
        for world in (1, 2, 3, 8):
            parts = [sharding.strong_shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert sharding.weak_shard(17_000_000, 3, 8) == (51_000_000, 17_000_000)
    with pytest.raises(ValueError):
        sharding.strong_shard(10, 2, 2)
    assert sharding.allreduce_result(5, 6) == (5, 6)                       # no group: identity
    red = sharding.ResultReducer()
    red.submit(7, 8)
    assert red.result() == (7, 8)


def test_bench_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--size", "small", "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip() == ""
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--size", "small", "--steps", "1", "--warmup", "0", "--lookups", "300000"],
                       env=env, capture_output=True, text=True)
    assert p.returncode == 0
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "lookups/s"
