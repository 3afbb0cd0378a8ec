#!/bin/bash
# bucket kernel variants on the quarter-size C4 instance: CTA shapes x bucket sizes
B=$PWD/genome-assembler_b200/build
for l1 in 8 9; do
  echo "== l1_bits=$l1"
  PROBE_L1=$l1 python scripts/bucket_probe.py 8192 2>&1 | tail -3
  for v in t384 t256 t128; do echo -n "$v "; GA_LIB=$B/libga_b200_$v.so PROBE_L1=$l1 python scripts/bucket_probe.py 8192 2>&1 | tail -1; done
  echo -n "r1 "; GA_LIB=$B/libga_b200_r1.so PROBE_L1=$l1 python scripts/bucket_probe.py 8192 2>&1 | tail -1
done
