timeout 200 python scripts/trace_front.py ico8 f 128 gpurun_out/trace_front_ico8_prof.npz 2>&1 | grep -E "^fwd|^bwd|local|dense"
