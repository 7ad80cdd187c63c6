#!/bin/bash
# N-GPU validation: multi-GPU tests + the bench line under torchrun (strong / bands extras)
set -u
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "not xxl" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f'gpurun_out/r02_bench_{sys.argv[1]}gpu.json').read().splitlines() if l.startswith('{')][-1])
    for k in ('value', 'ms_per_step', 'checksum_ok', 'n_gpus', 'e2e', 'strong', 'bands'):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print('no bench line', e)
PY
