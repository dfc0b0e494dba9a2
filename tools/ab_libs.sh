#!/bin/bash
# Dev tool: time the headline step with several builds of the library (same C ABI) on one box.
#   tools/ab_libs.sh build/libA.so build/libB.so ...
for rep in 1 2; do
for lib in "$@"; do
  echo "== $lib (rep $rep)"
  CVXPNPL_B200_LIB=$PWD/$lib python tools/seed_grace_sweep.py 40 ${SEEDS:-2} | grep -v "^grace"
done
done
