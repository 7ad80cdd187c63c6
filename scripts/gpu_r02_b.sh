#!/bin/bash
# Round 2, step b: the new parity tests (two live contexts, host-sample validation, canonical grid size,
# xsbench binary exit status, -b write/read) and the reworked bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_live or rejects_bad or canonical or xsbench_binary or golden_rows" 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
    for k in ('value', 'ms_per_step', 'checksum_ok', 'phase_ms', 'roofline', 'cpu_baseline', 'gpu_baseline', 'e2e', 'variants', 'strong', 'bands'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print('no bench line', e)
PY
