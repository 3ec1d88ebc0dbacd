# occupancy variants of the persistent kernel (run under gpurun)
run() { # variant, ring kb, workload, dtype, k
  if [ "$1" = main ]; then unset CHB_LIB_VARIANT; else export CHB_LIB_VARIANT=$1; fi
  echo "=== variant $1 ring $2 KB: $3 $4"
  CHB_RING_KB=$2 timeout 900 python scripts/sweep_experiments.py $3 $4 "$5" 2>&1 | grep -E "^k=|rror"
}
run main 36 cfg5 d 128,3
run d4 36 cfg5 d 128,3
run d5 27 cfg5 d 128,3
run d5 20 cfg5 d 128,3
unset CHB_LIB_VARIANT
echo "=== defaults"
timeout 900 python scripts/sweep_experiments.py ico8 f 128,64,32,3,1 2>&1 | grep -E "^k=|rror"
timeout 900 python scripts/sweep_experiments.py ico9 f 128 2>&1 | grep -E "^k=|rror"
timeout 900 python scripts/sweep_experiments.py ico6 f 128 2>&1 | grep -E "^k=|rror"
