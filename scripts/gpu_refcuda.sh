#!/bin/bash
# Reference cuda/ build (unmodified, sm_100a) on the same GPU: every -k variant and grid type.
set -u
for args in "-k 0" "-k 1" "-k 2" "-k 3" "-k 4" "-k 5" "-k 6" "-G hash -k 0" "-G nuclide -k 0" "-G hash -k 6"; do
  echo "== reference cuda: -m event -s large $args"
  timeout 300 oracle/_ref/XSBench_cuda_ref -m event -s large $args 2>&1 | grep -E "Runtime|Lookups/s|checksum" | tr '\n' ' '; echo
done
