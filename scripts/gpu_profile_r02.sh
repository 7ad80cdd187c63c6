#!/bin/bash
# Round-2 evidence: launch list of the bench command, full captures of the dominant kernel (xs_dense_kernel)
# and of the -k 0 kernel (xs_tile_kernel), the bench line with all its legs, the reference arm, other modes.
set -u
mkdir -p gpurun_out
T=r02
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-traffic-probe --no-extras > gpurun_out/${T}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xs_dense_kernel -s 1 -c 1 -f -o gpurun_out/${T}_dense \
    python scripts/quick_bench.py --kernels 6 --reps 1 > gpurun_out/${T}_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xs_tile_kernel -s 1 -c 1 -f -o gpurun_out/${T}_tile \
    python scripts/quick_bench.py --kernels 0 --reps 1 > gpurun_out/${T}_tile.log 2>&1
python bench.py --steps 20 --warmup 5 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
cut -c1-300 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/${T}_bench_reference_arm.json
cut -c1-300 gpurun_out/${T}_bench_reference_arm.json
echo "== other modes"
python scripts/quick_bench.py --method history --kernels 0 2>&1 | tail -1
python scripts/quick_bench.py --grid hash --kernels 0,4,6 2>&1 | tail -3
python scripts/quick_bench.py --grid nuclide --kernels 0,4,6 2>&1 | tail -3
python scripts/quick_bench.py --kernels 0,1,2,3,4,5,6 "" XSB200_ARITH=fused 2>&1 | tail -14
