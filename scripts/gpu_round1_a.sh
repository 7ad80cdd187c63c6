#!/bin/bash
# First contact with the B200: device facts, smoke, parity tests, reference CUDA baseline, bench.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_info.txt 2>&1
nproc >> $OUT/gpu_info.txt; free -g | head -2 >> $OUT/gpu_info.txt
python - >> $OUT/gpu_info.txt 2>&1 <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print(p)
print("L2", p.L2_cache_size, "sm", p.multi_processor_count, "smem/blk optin", getattr(p, "shared_memory_per_block_optin", None))
PY
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== reference cuda k0" ; timeout 300 oracle/_ref/XSBench_cuda_ref -m event -s large -k 0 2>&1 | grep -E "Runtime|Lookups/s|checksum" 
echo "== reference cuda k6" ; timeout 300 oracle/_ref/XSBench_cuda_ref -m event -s large -k 6 2>&1 | grep -E "Runtime|Lookups/s|checksum"
for g in 1 0; do
  echo "== bench gather=$g"; XSB200_GATHER=$g timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -2 | tee $OUT/bench_gather$g.json
done
echo "== driver k0..k6 (ours)"
for k in 0 1 2 3 4 5 6; do
  timeout 300 xsbench_b200/xsbench -m event -s large -k $k --reps 3 2>&1 | grep -E "Device time|Phases|Lookups/s|checksum" | tr '\n' ' '; echo
done
