#!/bin/sh
# A/B of two prebuilt libraries (libpmc_expA.so / libpmc_expB.so): correctness (smoke vs oracle) and timing
L=parallel-monte-carlo_b200
for t in A B A B; do
  cp $L/libpmc_exp$t.so $L/libpmc_b200.so
  echo "== $t"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  PMC_SWEEPS=300 python scripts/dev/quick16m.py
done
