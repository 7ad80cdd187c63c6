#!/bin/bash
# compute-sanitizer on what the third session of round 2 changed: sampler / locate kernels (digit counts, packed
# bucket entries), the sort's last pass, history mode (the sort counts its own digits there), host-sample call.
set -u
bash scripts/gpu_sanitize_k6.sh
X="xsbench_b200/xsbench -s small -g 300"
for args in "-m history -p 700 -l 9" "SORTED -m history -p 3000 -l 5" "-m event -l 20000 -k 4" "-m event -l 20000 -k 0"; do
  for tool in memcheck racecheck; do
    if [[ "$args" == SORTED* ]]; then args="${args#SORTED }"; export XSB200_HISTORY_SORTED=1; fi
    out=$(timeout 600 compute-sanitizer --tool $tool --print-limit 5 $X $args 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|Error|hazard" | head -4 | tr '\n' ' ')
    echo "[$tool] ${XSB200_HISTORY_SORTED:+SORTED }$args :: $out"
  done
  unset XSB200_HISTORY_SORTED
done
for tool in memcheck racecheck; do
  out=$(timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_lookup_samples_ragged_sizes or test_lookup_samples_rejects_bad_samples" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" | head -4 | tr '\n' ' ')
  echo "[$tool] host-sample call (pytest: ragged sizes, bad samples) :: $out"
done
