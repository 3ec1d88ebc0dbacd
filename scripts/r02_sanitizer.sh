#!/bin/bash
# compute-sanitizer over both sweep kernel families (ico5, k = 3 and k = 128): racecheck (shared-memory hazards of the
# rings), memcheck, synccheck.  Logs in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
for mode in panel front; do
  for k in 3 128; do
    for tool in racecheck memcheck synccheck; do
      CHB_SWEEP=$mode timeout 900 compute-sanitizer --tool $tool --log-file $O/r02_sanitizer_${mode}_k${k}_${tool}.log python scripts/solve_once.py ico5 f $k 2 > $O/r02_sanitizer_${mode}_k${k}_${tool}.out 2>&1
      echo "$mode k=$k $tool: $(tail -1 $O/r02_sanitizer_${mode}_k${k}_${tool}.log)"
    done
  done
done
