# chunked inverse-apply tasks, second pass (run under gpurun)
B="CHB_INV_CHUNK=8+CHB_INV_PANELS=256,CHB_INV_CHUNK=4+CHB_INV_PANELS=256,CHB_INV_CHUNK=8+CHB_INV_PANELS=512,CHB_INV_CHUNK=12+CHB_INV_PANELS=256,CHB_INV_CHUNK=16+CHB_INV_PANELS=512"
timeout 1500 python scripts/sweep_experiments.py ico9 f 128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
B="CHB_INV_CHUNK=8+CHB_INV_PANELS=256,CHB_INV_CHUNK=4+CHB_INV_PANELS=256,CHB_INV_CHUNK=12+CHB_INV_PANELS=256"
timeout 900 python scripts/sweep_experiments.py cfg5 d 128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
timeout 900 python scripts/sweep_experiments.py ico8 f 3,128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
