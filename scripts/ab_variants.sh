#!/bin/bash
# A/B of library variants on one box: scripts/ab_variants.sh "<variants>" <workload> <dtype> "<k list>"
for v in $1; do
  echo "=== variant ${v}"
  if [ "$v" = main ]; then unset CHB_LIB_VARIANT; else export CHB_LIB_VARIANT=$v; fi
  timeout 900 python scripts/sweep_experiments.py $2 $3 "$4" 2>&1 | grep -E "^k=|rror"
done
