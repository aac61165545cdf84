#!/bin/bash
# Does the parity harness notice a screening error bound that is too small?  Builds of the library with
# the bound scaled down (MCHB_SCREEN_ERR_SCALE, MCHB_SCREEN_SLACK) must FAIL tests/soak_parity.py:
#   a: bound 0, slack 0 (a decision is "certain" as soon as the screened value is on one side)
#   b: bound / 100, slack 1e-6
#   c: bound * -0.05 (claims certainty inside the true error band)
# and the product build must pass.  Run on a B200 box after building the variants:
#   nvcc ... -DMCHB_SCREEN_ERR_SCALE=0.0 -DMCHB_SCREEN_SLACK=0.0 -o mchap_b200/_lib/libmchap_b200_bound_a.so ...
for v in a b c; do
  export MCHB_LIB=$PWD/mchap_b200/_lib/libmchap_b200_bound_$v.so
  echo "== variant $v"
  python tests/soak_parity.py --items 3000 --steps 400 2>&1 | tail -2
  echo "exit code $?"
done
