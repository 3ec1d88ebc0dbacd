# inverse super-block length and ticket batching with 6 CTAs per SM (run under gpurun)
B="CHB_INV_PANELS=64,CHB_INV_PANELS=96,CHB_INV_PANELS=128,CHB_INV_PANELS=256,CHB_INV_PANELS=64+CHB_WIDE_LEVEL=65536,CHB_INV_PANELS=128+CHB_WIDE_LEVEL=65536+CHB_SPLIT_COST=6144+CHB_CHUNK_COST=3072,CHB_INV_PANELS=64+CHB_WIDE_LEVEL=1000000"
timeout 1500 python scripts/sweep_experiments.py ico9 f 128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
B="CHB_INV_PANELS=64,CHB_INV_PANELS=128,CHB_INV_PANELS=256,CHB_INV_PANELS=128+CHB_WIDE_LEVEL=65536"
timeout 900 python scripts/sweep_experiments.py cfg5 d 128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
timeout 900 python scripts/sweep_experiments.py ico8 f 3 0 "" "CHB_INV_PANELS=32,CHB_INV_PANELS=64,CHB_INV_PANELS=128,CHB_INV_PANELS=256" 2>&1 | grep -E "^k=|rror|^=="
