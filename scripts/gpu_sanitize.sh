#!/bin/bash
# compute-sanitizer over every kernel family at a tiny problem size.
set -u
X="xsbench_b200/xsbench -s small -g 300"
for args in "-m event -l 20000 -k 0" "-m event -l 20000 -k 2" "-m event -l 20000 -k 4" "-m event -l 20000 -k 5" "-m event -l 20000 -k 6" \
            "-m history -p 700 -l 9" "-m event -l 20000 -k 4 -G hash -h 100" "-m event -l 20000 -k 4 -G nuclide" "-m event -l 20000 -k 0 -G nuclide" \
            "-m event -l 20000 -k 4 --device-init" "-m event -l 20000 -k 6 --device-init" "-m event -l 20000 -k 6 -G hash -h 100" "-m event -l 20000 -k 6 -G nuclide" "BANDS -m event -l 20000 -k 6" \
            "DENSE -m event -l 20000 -k 6" "DENSE -m event -l 20000 -k 6 -G hash -h 100" "DENSE -m event -l 20000 -k 6 -G nuclide" "DENSE -m event -l 200000 -g 100 -k 6" \
            "FUSED -m event -l 200000 -g 100 -k 6" "-m event -l 20000 -k 0 -G hash -h 100" "-m event -l 20000 -k 1" "-m event -l 20000 -k 3"; do
  for tool in memcheck racecheck; do
    if [[ "$args" == BANDS* ]]; then args="${args#BANDS }"; export XSB200_BANDS=2 XSB200_BAND_INDEX=1; fi
    if [[ "$args" == DENSE* ]]; then args="${args#DENSE }"; export XSB200_DENSE_MIN=1; fi   # every material through xs_dense_kernel
    if [[ "$args" == FUSED* ]]; then args="${args#FUSED }"; export XSB200_DENSE_MIN=1 XSB200_ARITH=fused; fi
    out=$(timeout 600 compute-sanitizer --tool $tool --print-limit 5 $X $args 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|Error|hazard" | head -4 | tr '\n' ' ')
    echo "[$tool] ${XSB200_BANDS:+BANDS }${XSB200_DENSE_MIN:+DENSE }${XSB200_ARITH:+FUSED }$args :: $out"
  done
  unset XSB200_BANDS XSB200_BAND_INDEX XSB200_DENSE_MIN XSB200_ARITH
done
