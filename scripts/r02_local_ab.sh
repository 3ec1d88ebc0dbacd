# A/B of the local-task kernel (run under gpurun)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_pytest.log 2>&1; tail -3 gpurun_out/r02x_pytest.log
B="CHB_LOCAL_KERNEL=1,CHB_LOCAL_KERNEL_ROWS=20:36,CHB_LOCAL_KERNEL_ROWS=24:40,CHB_LOCAL_KERNEL_ROWS=36:60+CHB_LOCAL_KERNEL_BYTES=9000,CHB_LOCAL_KERNEL=0"
timeout 900 python scripts/sweep_experiments.py ico8 f 128,3 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
timeout 900 python scripts/sweep_experiments.py cfg5 d 128 0 "" "$B" 2>&1 | grep -E "^k=|rror|^=="
timeout 900 python scripts/sweep_experiments.py ico9 f 128 0 "" "CHB_LOCAL_KERNEL=1,CHB_LOCAL_KERNEL=0" 2>&1 | grep -E "^k=|rror|^=="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chb_ --csv --log-file gpurun_out/r02w_launches_ico8.csv python scripts/solve_once.py ico8 f 128 3 > /dev/null 2>&1
grep -E "chb_(local|sweep)" gpurun_out/r02w_launches_ico8.csv | cut -d, -f5,9,15 | tail -4
