#!/bin/bash
# Round-1 evidence: launch list of the bench command, full capture of the dominant kernel, a 2-GPU bench.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_bench_under_ncu.log 2>&1
# dominant kernel: a mid fuel window (launch #6 of the window kernel in a step); skip init + warm-up step
ncu --set full --clock-control none --import-source on -k regex:xs_window_kernel -s 17 -c 1 -f -o gpurun_out/r01b_window \
    python scripts/quick_bench.py --kernels 4 --reps 1 > gpurun_out/r01b_window.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xs_event_kernel -s 1 -c 1 -f -o gpurun_out/r01b_event_k0 \
    python scripts/quick_bench.py --kernels 0 --reps 1 > gpurun_out/r01b_event.log 2>&1
python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r01b_bench.json
cat gpurun_out/r01b_bench.json | cut -c1-400
echo "== history mode + other grids"
python scripts/quick_bench.py --method history --kernels 0 2>&1 | tail -1
python scripts/quick_bench.py --grid hash --kernels 0,4 2>&1 | tail -2
python scripts/quick_bench.py --grid nuclide --kernels 0,4 2>&1 | tail -2
