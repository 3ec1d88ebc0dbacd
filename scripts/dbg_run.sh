for d in 0 1 2 4 7 63; do echo "== CHB_FRONT_DEBUG=$d"; CHB_FRONT_DEBUG=$d timeout 200 python scripts/trace_front.py ico8 f 128 2>&1 | grep -E "^fwd|^bwd|local:|dense:"; done
