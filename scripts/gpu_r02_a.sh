#!/bin/bash
# Round 2, first contact: baseline bench line of the round-1 build on this pod, and the reference
# cuda/ build (unmodified, sm_100a) device-timed per kernel (ncu gpu__time_duration) for -k 0 / -k 6.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader > gpurun_out/r02a_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.json
for k in 0 6; do
  echo "== reference cuda -k $k (its own host timer)"
  timeout 300 oracle/_ref/XSBench_cuda_ref -m event -s large -k $k 2>&1 | grep -E "Runtime|Lookups/s|checksum" | tr '\n' ' '; echo
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file gpurun_out/r02a_refcuda_k${k}_launches.csv oracle/_ref/XSBench_cuda_ref -m event -s large -k $k > gpurun_out/r02a_refcuda_k${k}.log 2>&1
  python scripts/sum_launches.py gpurun_out/r02a_refcuda_k${k}_launches.csv > gpurun_out/r02a_refcuda_k${k}_summary.md; cat gpurun_out/r02a_refcuda_k${k}_summary.md
done
