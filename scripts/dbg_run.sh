timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0d5d931 ""; do echo "== variant $v"; CHB_LIB_VARIANT=$v timeout 300 python scripts/sweep_experiments.py ico8 f 128,3 0 "" "CHB_FRONT_WARPS=7" 2>&1 | grep -E "k=|Error|error"; done
