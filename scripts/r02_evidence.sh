#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): bench lines, ncu launch list + full capture of the benched build,
# reference GPU kernels, compute-sanitizer logs.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r02z_clocks.csv 2>/dev/null &
SMI=$!
# 1. headline: cfg4 (default flags), then cfg2 / cfg5 stand-in / cfg3 stand-in without the CPU baseline leg
timeout 1500 python bench.py > $O/r02z_bench_cfg4.json 2> $O/r02z_bench_cfg4.err
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > $O/r02z_bench_cfg2.json 2> $O/r02z_bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > $O/r02z_bench_cfg5.json 2> $O/r02z_bench_cfg5.err
CHB_CFG3_SIDE=48 timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $O/r02z_bench_cfg3_48.json 2> $O/r02z_bench_cfg3_48.err
# 2. front image on the same workloads (alternative path)
CHB_SWEEP=front timeout 900 python bench.py --no-cpu-baseline > $O/r02z_bench_cfg4_front.json 2> $O/r02z_bench_cfg4_front.err
# 3. ncu: launch list of a bench run, full capture of the two sweeps of the benched build (cfg4)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02z_launches_cfg4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r02z_bench_under_ncu.json 2>/dev/null
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:chb_sweep -s 4 -c 2 -o $O/r02z_sweep_cfg4 python scripts/solve_once.py cfg4 f 128 3 > $O/r02z_ncu_cfg4.log 2>&1
# 4. the reference's own GPU kernels on this B200
timeout 600 python bench.py --impl reference-gpu --workload cfg2 --steps 5 --warmup 1 > $O/r02z_refgpu_cfg2.json 2> $O/r02z_refgpu_cfg2.err
timeout 900 python bench.py --impl reference-gpu --workload cfg4 --steps 3 --warmup 1 > $O/r02z_refgpu_cfg4.json 2> $O/r02z_refgpu_cfg4.err
kill $SMI 2>/dev/null
