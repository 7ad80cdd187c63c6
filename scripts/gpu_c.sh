#!/bin/bash
set -u
python -m pytest tests -m "gpu and not slow" -x -q 2>&1 | tail -3
python scripts/quick_bench.py --kernels 4,5,6 2>&1 | tail -3
for c in 0 3; do
  XSB200_E2E_CHUNKS=$c python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > /tmp/b.json
  python - <<PY
import json
d=json.load(open('/tmp/b.json'))
print("chunks", $c, "value %.1f M/s" % (d["value"]/1e6), "e2e %.1f M/s" % (d["e2e"]["value"]/1e6), d["checksum_ok"], d["e2e"]["checksum_matches_device_sampled"])
PY
done
