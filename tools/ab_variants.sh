#!/bin/bash
# Time the Griffin-Lim iteration kernel for every library build under build_variants/ (kernel A/B experiments).
for so in build_variants/*.so; do
  echo "== $so"
  S2ST_B200_LIB=$PWD/$so timeout 300 python tools/time_pass.py 0 2>&1 | tail -1
  [ -n "$SMALL" ] && S2ST_B200_LIB=$PWD/$so timeout 300 python tools/time_small.py 2>&1 | tail -3
done
