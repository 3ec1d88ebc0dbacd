for v in 6505e67 nopipe ""; do echo "== variant $v"; CHB_LIB_VARIANT=$v timeout 300 python scripts/sweep_experiments.py ico8 f 128,3 0 "" "CHB_FRONT_WARPS=7" 2>&1 | grep -E "k=|Error|error"; done
