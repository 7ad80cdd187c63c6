#!/bin/bash
# the whole -m gpu suite (what the driver runs), smoke(), and the bench line with all its legs
set -u
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench.json').read().splitlines() if l.startswith('{')][-1])
    for k in ('value', 'ms_per_step', 'checksum_ok', 'phase_ms', 'cpu_baseline', 'e2e', 'variants', 'strong', 'gpu_launches'):
        print(k, json.dumps(d.get(k))[:600])
    r = d['roofline']; print('roofline', {k: r[k] for k in ('bound', 'achieved', 'peak', 'frac', 'floor_ms', 'kernel_ms', 'traffic')}, r['hbm'])
    print('gpu_baseline', json.dumps({k: d['gpu_baseline'].get(k) for k in ('k0', 'k6')}))
except Exception as e:
    print('no bench line', e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
