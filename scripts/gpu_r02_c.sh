#!/bin/bash
# Round 2, step c: the rewritten xs_dense_kernel -- parity first, then timing (exact and fused arithmetic).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or sorted_pipeline or grid_points or canonical or division or sweep_path or two_live or history or large_fuel" 2>&1 | tail -8
for arith in exact fused; do
  XSB200_ARITH=$arith timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-extras --no-traffic-probe > gpurun_out/r02c_bench_$arith.json 2> gpurun_out/r02c_bench_$arith.err; echo "bench $arith rc=$?"
  python - $arith <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/r02c_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], d['checksum_ok'], d['phase_ms'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'])
except Exception as e:
    print('no bench line', e); print(open(f'gpurun_out/r02c_bench_{sys.argv[1]}.err').read()[-1500:])
PY
done
