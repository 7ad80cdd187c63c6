set -u
X="xsbench_b200/xsbench -s small -g 300"
for args in "-m event -l 20000 -k 6" "-m event -l 20000 -k 6 -G hash -h 100" "DENSE -m event -l 20000 -k 6" "DENSE -m event -l 20000 -k 6 -G hash -h 100" "DENSE -m event -l 20000 -k 6 -G nuclide" "DENSE -m event -l 200000 -g 100 -k 6"; do
  for tool in memcheck racecheck; do
    if [[ "$args" == DENSE* ]]; then args="${args#DENSE }"; export XSB200_DENSE_MIN=1; fi
    out=$(timeout 600 compute-sanitizer --tool $tool --print-limit 5 $X $args 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|Error|hazard" | head -4 | tr '\n' ' ')
    echo "[$tool] ${XSB200_DENSE_MIN:+DENSE }$args :: $out"
  done
  unset XSB200_DENSE_MIN
done
